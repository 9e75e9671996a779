#!/bin/bash
for v in "" "-DFFN_EXP_NOCONV" "-DFFN_EXP_NOFIX" "-DFFN_EXP_NOPROD" "-DFFN_EXP_NOCONV -DFFN_EXP_NOFIX -DFFN_EXP_NOPROD"; do
  GNB_EXTRA_NVCC_FLAGS="$v" python graphnets.jl_b200/build.py --force > /dev/null 2>&1
  TAG="[$v]" timeout 200 python scratch/ffn_probe.py tc_ffn256 tc_linear 2>&1 | tail -1
done
python graphnets.jl_b200/build.py --force > /dev/null 2>&1
