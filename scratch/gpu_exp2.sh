#!/bin/bash
# $1 = extra defines; prints bench A/B and the timing means
GNB_EXTRA_NVCC_FLAGS="$1" python graphnets.jl_b200/build.py --force > /dev/null 2>&1
echo "=== build flags: $1"
timeout 100 python -m pytest tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -1
bash scratch/gpu_ab.sh
GNB_EXTRA_NVCC_FLAGS="-DGNB_TC_TIMING $1" python graphnets.jl_b200/build.py --force > /dev/null 2>&1
timeout 120 python scratch/tc_timing.py > gpurun_out/tc_timing_exp.log 2>&1
grep "mean" gpurun_out/tc_timing_exp.log
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
