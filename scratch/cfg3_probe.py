"""BASELINE config 3 (examples/sort model: (0,100,0) -> 384 x3, 2 cores, -> (2,2,0); 4096 fully connected graphs of 8-64 nodes,
padded vector-mode batch path) at full size: batch (lowering) time and forward time, auto vs fp32 precision."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
import graphnets_b200 as gn, workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = W.make_workload("cfg3", B=B)
layers = W.model_params("cfg3")
model = W.to_gn_model(gn, layers)
t0 = time.perf_counter(); x = gn.batch(W.as_batch_input(w)); torch.cuda.synchronize(); t_batch = time.perf_counter() - t0
g = x.graphs
print("cfg3 B=%d: E=%d N=%d PN=%d  batch() %.1f ms (host python + H2D + GPU lowering)" % (B, g.E, g.N, g.PN if hasattr(g, "PN") else -1, t_batch * 1e3))
eng = g.engine
for prec, steps in (("auto", 3), ("fp32", 1)):
    for _ in range(2): y = model(x, precision=prec)
    torch.cuda.synchronize()
    eng.set_profiling(True); eng.read_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): y = model(x, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = eng.read_profile(); eng.set_profiling(False)
    top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:4]
    print("  %s: %.1f ms/forward = %.3g edges/s, %.3g graphs/s; top: %s" % (prec, ms, g.E / ms * 1e3, B / ms * 1e3,
          ", ".join("%s %.1f ms" % (k, v["ms"] / steps) for k, v in top)))
