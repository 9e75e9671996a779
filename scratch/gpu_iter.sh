#!/bin/bash
# edit-compile-measure loop: tensor-path parity tests, per-kernel timing, clock64 phase timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_reference_suite.py -m gpu -x -q 2>&1 | tail -4
TAG=prod timeout 120 python scratch/edge_probe.py tc_edge_core tc_node_core tc_agg graph_post 2>&1 | tail -1
if [ -n "$TIMING" ]; then
GNB_EXTRA_NVCC_FLAGS=-DGNB_TC_TIMING python graphnets.jl_b200/build.py --force > /dev/null 2>&1
timeout 120 python scratch/tc_timing.py > gpurun_out/tc_timing.log 2>&1
grep -A5 "==\|mean" gpurun_out/tc_timing.log | grep -v "^--"
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
fi
