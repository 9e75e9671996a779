#!/bin/bash
# timing experiment builds: $1 = extra nvcc defines
GNB_EXTRA_NVCC_FLAGS="-DGNB_TC_TIMING $1" python graphnets.jl_b200/build.py --force > /dev/null 2>&1
timeout 120 python scratch/tc_timing.py > gpurun_out/tc_timing_exp.log 2>&1
echo "=== $1"; grep "mean" gpurun_out/tc_timing_exp.log
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
