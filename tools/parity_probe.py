"""Error of the tensor path against the float64 oracle after 1..4 cores of the cfg4 model (no decoder), 512 graphs:
shows where along the stack the bf16 error enters.   [GNB_LIB_VARIANT=...] python tools/parity_probe.py [graphs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402
from oracle import gn_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = W.make_workload("cfg4", B=B)
layers_all = W.model_params("cfg4")
g = O.lower(W.adj_list(w))
ef, nf, gf = W.compact_inputs(w)
x = gn.batch(W.as_batch_input(w))
for depth in (1, 2, 3, 4, 5):
    layers = layers_all[:1 + depth] if depth <= 4 else layers_all
    ref = O.forward_sparse(layers, g, ef, nf, gf)
    model = W.to_gn_model(gn, layers)
    y = model(x, precision="auto")
    x.graphs.engine.sync()
    got = [t.compact.cpu().numpy() for t in (y.ef, y.nf, y.gf)]
    print("enc + %d cores%s: " % (min(depth, 4), " + dec" if depth == 5 else "") +
          "  ".join("%s max %.2e rms %.2e" % (n, O.rel_err(a, b), O.rms_err(a, b)) for n, a, b in zip(("ef", "nf", "gf"), got, ref)), flush=True)
