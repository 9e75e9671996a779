"""Times the cfg4 forward (device-resident) and prints per-kernel ms; used for perf experiments (GNB_EDGE_FLAGS etc.)."""
import sys, os; sys.path.insert(0, '.')
import torch
import graphnets_b200 as gn, workloads as W
from bench import synth
adj, ef, nf = synth("cfg5", 4096, 1000)
model = W.to_gn_model(gn, W.model_params("cfg5"))
x = gn.batch_compact(adj, ef, nf)
eng = x.graphs.engine
for _ in range(3): y = model(x, precision="auto")
torch.cuda.synchronize()
eng.set_profiling(True); eng.read_profile()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record()
for _ in range(steps): y = model(x, precision="auto")
e1.record(); torch.cuda.synchronize()
prof = eng.read_profile()
ms = e0.elapsed_time(e1) / steps
keys = sys.argv[1:] or ["tc_edge_core", "tc_node_core"]
print("%s ms/step %.3f  " % (os.environ.get("TAG", ""), ms) + "  ".join("%s %.3f" % (k, prof[k]["ms"] / prof[k]["launches"]) for k in keys if k in prof))
