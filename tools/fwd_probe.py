"""Runs a few forwards of a BASELINE config (for ncu captures):  python tools/fwd_probe.py [cfg4] [graphs] [precision] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402
from bench import synth          # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prec = sys.argv[3] if len(sys.argv) > 3 else "auto"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
adj, ef, nf = synth(cfg, B, 1000)
model = W.to_gn_model(gn, W.model_params(cfg))
x = gn.batch_compact(adj, ef, nf)
for _ in range(iters):
    y = model(x, precision=prec)
x.graphs.engine.sync()
print("done", x.graphs.E, x.graphs.N, x.graphs.engine.launches)
