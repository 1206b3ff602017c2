#!/bin/bash
set -x
mkdir -p gpurun_out
tools/microbench/fp32_pipes > gpurun_out/r2l4_fp32_pipes.txt 2>&1
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=4
{
echo "scalar term group=16"; NICP_GROUP=16 NICP_GROUP_MINB=15 python tools/tune_corr.py
echo "packed term group=16"; NICP_GROUP=16 NICP_GROUP_MINB=16 python tools/tune_corr.py
} > gpurun_out/r2l4_tune.txt 2>&1
export TUNE_REPS=1
NICP_GROUP=16 NICP_GROUP_MINB=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corr_lin_group -s 12 -c 1 \
  -o gpurun_out/r2l4_group python tools/tune_corr.py > gpurun_out/r2l4_ncu.log 2>&1
