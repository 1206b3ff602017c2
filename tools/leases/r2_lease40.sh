#!/bin/bash
mkdir -p gpurun_out
export NICP_BATCH_SLOTS=256 TUNE_REPS=5
for shape in "2 128" "4 64" "8 32" "16 16"; do
  set -- $shape
  echo "currents=$1 candidates=$2"; TUNE_CUR=$1 TUNE_CAND=$2 timeout 300 python tools/tune_corr.py | tail -1
done > gpurun_out/r2l40_shapes.txt 2>&1
cat gpurun_out/r2l40_shapes.txt
