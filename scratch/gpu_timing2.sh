#!/bin/bash
# clock64 phase timing of the edge kernel for several GNB_EDGE_FLAGS values
mkdir -p gpurun_out
GNB_EXTRA_NVCC_FLAGS=-DGNB_TC_TIMING python graphnets.jl_b200/build.py --force > /dev/null 2>&1
for f in ${FLAGS:-0 127}; do
  echo "######## flags=$f"
  GNB_EDGE_FLAGS=$f timeout 120 python scratch/tc_timing.py > gpurun_out/tc_timing_f$f.log 2>&1
  grep -A3 "==\|mean" gpurun_out/tc_timing_f$f.log | grep -v "cta  77" | grep -v "^--"
done
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
