"""Per-kernel time of one forward of a BASELINE config (library profile, CUDA events around every launch):
    python tools/cfg_profile.py cfg2 1024 fp32        python tools/cfg_profile.py cfg5 4096 auto"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                     # noqa: E402
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else None
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
os.environ.setdefault("GNB_PROFILE_SHAPES", "1")
w = W.make_workload(cfg, B=B)
model = W.to_gn_model(gn, W.model_params(cfg))
x = gn.batch(W.as_batch_input(w))
eng = x.graphs.engine
for _ in range(3):
    y = model(x, precision=prec)
eng.sync()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
ev0.record()
for _ in range(n):
    y = model(x, precision=prec)
ev1.record()
torch.cuda.synchronize()
print("%s B=%s %s: E=%d N=%d, forward %.3f ms" % (cfg, B, prec, x.graphs.E, x.graphs.N, ev0.elapsed_time(ev1) / n))
eng.set_profiling(True)
eng.read_profile()
for _ in range(5):
    y = model(x, precision=prec)
prof = eng.read_profile()
tot = sum(v["ms"] for v in prof.values()) / 5
print("  sum of kernels %.3f ms per forward" % tot)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print("  %-28s %3d launches/forward  %8.1f us/launch  %6.1f %%" % (k, v["launches"] // 5, 1e3 * v["ms"] / v["launches"], 100 * v["ms"] / 5 / tot))
