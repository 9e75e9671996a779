"""Which core-output tensor's bf16 error does the (fp32) decoder amplify?  Runs enc + 4 cores of cfg4 on the GPU ("auto"), then
the float64 oracle decoder on mixtures of GPU / oracle core outputs.   python tools/dec_sensitivity.py [graphs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402
from oracle import gn_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = W.make_workload("cfg4", B=B)
layers = W.model_params("cfg4")
g = O.lower(W.adj_list(w))
ef, nf, gf = W.compact_inputs(w)
x = gn.batch(W.as_batch_input(w))
ref_core = O.forward_sparse(layers[:-1], g, ef, nf, gf)
y = W.to_gn_model(gn, layers[:-1])(x, precision="auto")
x.graphs.engine.sync()
got = [t.compact.cpu().numpy().astype(np.float64) for t in (y.ef, y.nf, y.gf)]
print("core outputs: " + "  ".join("%s max %.2e rms %.2e" % (n, O.rel_err(a, b), O.rms_err(a, b)) for n, a, b in zip("eng", got, ref_core)))
dec = layers[-1][1]
ref = O.gnblock_sparse(dec, g, *ref_core)
for label, mix in (("all gpu", got), ("only ef gpu", [got[0], ref_core[1], ref_core[2]]), ("only nf gpu", [ref_core[0], got[1], ref_core[2]]),
                   ("only gf gpu", [ref_core[0], ref_core[1], got[2]])):
    out = O.gnblock_sparse(dec, g, *mix)
    print("%-12s decoder out: " % label + "  ".join("%s max %.2e rms %.2e" % (n, O.rel_err(a, b), O.rms_err(a, b)) for n, a, b in zip("eng", out, ref)))
# error structure of the edge features: how much of it is common to all edges (coherent) ?
d = got[0] - ref_core[0]
print("ef error: rms %.3e, rms of its mean over edges %.3e, rms of ref %.3e" % (np.sqrt((d ** 2).mean()), np.sqrt((d.mean(0) ** 2).mean()), np.sqrt((ref_core[0] ** 2).mean())))
d = got[2] - ref_core[2]
print("gf error: rms %.3e, rms of its mean over graphs %.3e, rms of ref %.3e" % (np.sqrt((d ** 2).mean()), np.sqrt((d.mean(0) ** 2).mean()), np.sqrt((ref_core[2] ** 2).mean())))
