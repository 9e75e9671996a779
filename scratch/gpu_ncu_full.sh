#!/bin/bash
# one --set full capture of the hot kernels of a steady-state bench step (k_edge5 edge + node launches first)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_edge5|k_tc_proj|k_graph_post|k_wide|k_narrow2' --launch-skip 30 -c 16 -f -o gpurun_out/r01_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/
