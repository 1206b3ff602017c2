#!/bin/bash
mkdir -p gpurun_out
export NICP_BATCH_SLOTS=256 TUNE_REPS=5
for cfg in "4 64 16" "4 64 32" "2 128 16" "2 128 32" "2 128 26"; do
  set -- $cfg
  echo "currents=$1 candidates=$2 NICP_GROUP=$3"; TUNE_CUR=$1 TUNE_CAND=$2 NICP_GROUP=$3 timeout 300 python tools/tune_corr.py | tail -1
done > gpurun_out/r2l43_tune.txt 2>&1
cat gpurun_out/r2l43_tune.txt
NICP_GROUP=32 timeout 900 python -m pytest tests -m gpu -q -x -k "determinism or batch or grouped or priors or epoch or sharded or loop_closure or random_scenes" 2>&1 | tail -2
