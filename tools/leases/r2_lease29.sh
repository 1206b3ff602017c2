#!/bin/bash
mkdir -p gpurun_out
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
for g in 16 32 16 32 24; do
  echo "NICP_GROUP=$g"; NICP_GROUP=$g timeout 300 python tools/tune_corr.py | tail -1
done > gpurun_out/r2l29_tune.txt 2>&1
cat gpurun_out/r2l29_tune.txt
