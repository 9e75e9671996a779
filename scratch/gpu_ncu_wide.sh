#!/bin/bash
# --set full capture of the wide-core kernels (fused FFN-256 on the edge rows, generic linear layer) in a cfg5-shape forward
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_ffn|k_tc_lin' --launch-skip 60 -c 4 -f -o gpurun_out/r01_wide_full python scratch/ffn_probe.py tc_ffn256 tc_linear > gpurun_out/ncu_wide.log 2>&1
tail -3 gpurun_out/ncu_wide.log | cut -c1-200
ls -la gpurun_out/r01_wide_full.ncu-rep
