#!/bin/bash
# A/B of build-flag variants: prints per-kernel time and the phase timing means for each
mkdir -p gpurun_out
for v in "${@}"; do
  echo "######## variant: $v"
  GNB_EXTRA_NVCC_FLAGS="$v" python graphnets.jl_b200/build.py --force > /dev/null 2>&1
  TAG=time timeout 120 python scratch/edge_probe.py tc_edge_core tc_node_core 2>&1 | tail -1
  GNB_EXTRA_NVCC_FLAGS="-DGNB_TC_TIMING $v" python graphnets.jl_b200/build.py --force > /dev/null 2>&1
  timeout 120 python scratch/tc_timing.py 2>&1 | grep -v "cta   0" | grep -A12 "mean\|weights"  | grep -v "^--\|^=="
done
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
