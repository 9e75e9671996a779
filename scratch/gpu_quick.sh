#!/bin/bash
# quick GPU check: parity tests for the tensor path + bench line
set -x
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("ms_per_step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "edges/s %.3g"%d["value"])
for k,v in d["kernels"].items(): print("  %-24s %8.3f ms/step  %5.1f launches"%(k,v["ms_per_step"],v["launches_per_step"]))
print(d["roofline"])
PY
