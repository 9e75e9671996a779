"""graphnets.jl_b200 - B200-native GNBlock / GNCore forward path behind GraphNets.jl's API.

Exports mirror the reference module (src/GraphNets.jl:12-50).  The directory name carries a dot,
so import it through the repo-root shim:  `import graphnets_b200 as gn`.
"""
from ._lib import lib, LIB_PATH, GnbError
from .engine import get_engine
from .api import (GNData, GNGraphBatch, Padded, batch, batch_compact, batch_coo, pack_adjacency_bits, unbatch, checks, efview, nfview, gfview,
                  flatunpaddednf, flatunpaddedef, logitcrossentropy, collapsef, unpaddedcollapsedef, flatunpaddedcollapsedef)
from .layers import (GNBlock, GNCore, GNCoreList, GNSequential, GNFeedForward, GNGraphNorm, Dense, Chain,
                     LayerNorm, Dropout, set_precision, get_precision)
from .shard import shard_ranges, shard_batch
from .train import Trainer

__all__ = [
    "GNGraphBatch", "batch", "batch_compact", "batch_coo", "pack_adjacency_bits", "unbatch", "GNBlock", "GNCore", "GNCoreList", "GNSequential", "efview", "nfview",
    "gfview", "flatunpaddednf", "flatunpaddedef", "logitcrossentropy", "collapsef", "unpaddedcollapsedef", "flatunpaddedcollapsedef",
    "GNData", "Padded", "set_precision", "get_precision", "get_engine", "shard_ranges", "shard_batch", "Trainer",
]
