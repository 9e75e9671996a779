#!/bin/bash
# --set full capture of one steady-state launch of the fused kernel on the edges (E rows) and one on the nodes;
# then a lighter capture (no source counters) of the other hot kernels
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edge5' --launch-skip 6 -c 2 -f -o gpurun_out/r01_edge5_full python scratch/edge_probe.py > gpurun_out/ncu_edge.log 2>&1
tail -3 gpurun_out/ncu_edge.log | cut -c1-200
timeout 300 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -k regex:'k_tc_proj|k_graph_post|k_wide|k_narrow2|k_graph_pre|k_zsum' --launch-skip 40 -c 14 -f -o gpurun_out/r01_others python scratch/edge_probe.py > gpurun_out/ncu_others.log 2>&1
tail -3 gpurun_out/ncu_others.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
