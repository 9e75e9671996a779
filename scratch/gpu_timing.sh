#!/bin/bash
# rebuild with clock64 stamps, run the phase timing, rebuild the production library
GNB_EXTRA_NVCC_FLAGS=-DGNB_TC_TIMING python graphnets.jl_b200/build.py --force > /dev/null 2>&1
timeout 120 python scratch/tc_timing.py > gpurun_out/tc_timing.log 2>&1
grep -A3 "==\|mean" gpurun_out/tc_timing.log | grep -v "cta  77" | grep -v "^--"
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
