#!/bin/bash
# full check: gpu tests, bench line, ncu launch list of the bench command
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"], "edges/s %.3g"%d["value"])
for k,v in d["kernels"].items(): print("  %-24s %8.3f ms/step  %5.1f launches"%(k,v["ms_per_step"],v["launches_per_step"]))
print(d["roofline"]); print(d.get("cpu_baseline")); print(d.get("clocks"))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log | cut -c1-300
