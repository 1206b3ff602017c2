#!/bin/bash
mkdir -p gpurun_out
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
for v in 1 2 1 2; do
  echo "NICP_CORR_VARIANT=$v"; NICP_CORR_VARIANT=$v timeout 300 python tools/tune_corr.py | tail -1
done > gpurun_out/r2l37_tune.txt 2>&1
cat gpurun_out/r2l37_tune.txt
