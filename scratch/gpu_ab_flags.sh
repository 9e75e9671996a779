#!/bin/bash
# A/B of nvcc define sets: per-kernel times for each ("$@" = one quoted flag set per variant)
for v in "${@}"; do
  GNB_EXTRA_NVCC_FLAGS="$v" python graphnets.jl_b200/build.py --force > /dev/null 2>&1 || { echo "[$v] BUILD FAILED"; continue; }
  TAG="[$v]" timeout 120 python scratch/edge_probe.py ${KEYS:-tc_edge_core graph_post} 2>&1 | tail -1
done
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
