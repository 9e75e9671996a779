#!/bin/bash
# A/B of library variants on the cfg4 forward: per-kernel ms from the library profile (tools/edge_timing.py prints it last)
for v in "" "$@"; do
  echo "=== variant '${v}'"
  GNB_LIB_VARIANT=$v python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, graphnets_b200 as gn, workloads as W
from bench import synth
adj, ef, nf = synth("cfg4", 4096, 1000)
model = W.to_gn_model(gn, W.model_params("cfg4"))
x = gn.batch_compact(adj, ef, nf)
eng = x.graphs.engine
for _ in range(3): y = model(x, precision="auto")
eng.sync()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10): y = model(x, precision="auto")
ev1.record(); torch.cuda.synchronize()
print("forward %.3f ms" % (ev0.elapsed_time(ev1) / 10))
eng.set_profiling(True); eng.read_profile()
for _ in range(5): y = model(x, precision="auto")
prof = eng.read_profile()
print("  " + "  ".join("%s %.3f" % (k, v["ms"] / v["launches"]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]) + "   (ms/launch)")
PY
done
