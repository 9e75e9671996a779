"""Multi-GPU sharding of a graph batch.

Graphs in a batch never interact (every reference broadcaster is block-diagonal per graph,
src/gngraphbatch.jl:136-211) and weights are replicated, so the forward shards over contiguous
graph ranges with NO data-path collective.  Ranges are balanced by EDGE count (the edge update
dominates), computed from host-side counts only - this module is pure host logic and is covered
by world_size-2 gloo tests on CPU."""
import numpy as np


def shard_ranges(edge_counts, world_size):
    """Contiguous graph ranges [(lo, hi)] * world_size with near-equal edge totals.

    Greedy prefix split: rank r ends at the first graph where the running edge count reaches
    (r+1)/world_size of the total.  Deterministic and identical on every rank."""
    m = np.asarray(edge_counts, np.int64)
    B = m.size
    if world_size <= 1:
        return [(0, B)]
    pref = np.concatenate([[0], np.cumsum(m)])
    total = pref[-1]
    bounds = [0]
    for r in range(1, world_size):
        if total == 0:
            cut = (B * r) // world_size
        else:
            cut = int(np.searchsorted(pref, total * r / world_size, side="left"))
        cut = min(max(cut, bounds[-1]), B)
        bounds.append(cut)
    bounds.append(B)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def shard_batch(x, rank, world_size):
    """Slice an UNBATCHED vector-mode input {graphs, ef, nf, gf} to this rank's graph range."""
    graphs = x["graphs"]
    counts = [int(np.count_nonzero(np.asarray(a) == 1)) for a in graphs]
    lo, hi = shard_ranges(counts, world_size)[rank]
    cut = lambda v: None if v is None else v[lo:hi]
    return dict(graphs=graphs[lo:hi], ef=cut(x["ef"]), nf=cut(x["nf"]), gf=cut(x["gf"])), (lo, hi)
