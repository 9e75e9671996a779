#!/bin/bash
# cfg5 shape (hidden 256) on one GPU: generic tcgen05 linear path vs the fp32 CUDA-core path
mkdir -p gpurun_out
for p in auto fp32; do
  timeout 300 python bench.py --config cfg5 --graphs ${GRAPHS:-4096} --steps 3 --warmup 3 --precision $p --no-cpu-baseline > gpurun_out/bench_cfg5_$p.json 2> gpurun_out/bench_cfg5_$p.err || tail -3 gpurun_out/bench_cfg5_$p.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg5_$p.json'))
print("$p: ms_per_step %.2f edges/s %.3g e2e ms %.1f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
for k,v in list(d["kernels"].items())[:8]: print("   %-24s %9.3f ms/step %6.1f launches"%(k,v["ms_per_step"],v["launches_per_step"]))
print("   roofline", {k:d["roofline"][k] for k in ("kernel","achieved","unit","frac")}, "model", d["model_roofline"])
PY
done
