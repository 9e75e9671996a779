"""fp32 path timing: cfg4 forward at a given number of graphs + config 5's training step (k_linear dominated)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                     # noqa: E402
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402
import bench                     # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
adj, ef, nf = bench.synth("cfg4", B, 1000)
model = W.to_gn_model(gn, W.model_params("cfg4"))
x = gn.batch_compact(adj, ef, nf)
for _ in range(2):
    y = model(x, precision="fp32")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    y = model(x, precision="fp32")
e1.record()
torch.cuda.synchronize()
print("cfg4 fp32 forward, %d graphs: %.2f ms" % (B, e0.elapsed_time(e1) / 3))
r = bench.train_leg(torch, gn, W, None, 1, 256)
print("cfg5 training step, 256 graphs: %.1f ms" % r["ms_per_step"])
