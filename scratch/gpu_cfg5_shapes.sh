#!/bin/bash
GNB_PROFILE_SHAPES=1 timeout 300 python bench.py --config cfg5 --graphs ${GRAPHS:-4096} --steps 3 --warmup 3 --precision auto --no-cpu-baseline > gpurun_out/bench_cfg5_shapes.json 2> gpurun_out/bench_cfg5_shapes.err || tail -3 gpurun_out/bench_cfg5_shapes.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg5_shapes.json'))
print("ms_per_step %.2f"%d["ms_per_step"])
for k,v in list(d["kernels"].items())[:14]: print("   %-28s %9.3f ms/step %6.1f launches  %.3f ms/launch"%(k,v["ms_per_step"],v["launches_per_step"],v["ms_per_step"]/v["launches_per_step"]))
PY
