#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "determinism or batch or grouped or priors or epoch or sharded or loop_closure" > gpurun_out/r2l25_pytest.log 2>&1
tail -3 gpurun_out/r2l25_pytest.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
for cfg in "0 1" "14 1" "13 1" "12 1" "12 2" "11 2" "10 2"; do
  set -- $cfg
  echo "NICP_OVERLAP_WALK_CTAS=$1 NICP_OVERLAP_PROJ_CTAS=$2"; NICP_OVERLAP_WALK_CTAS=$1 NICP_OVERLAP_PROJ_CTAS=$2 timeout 300 python tools/tune_corr.py | tail -1
done > gpurun_out/r2l25_tune.txt 2>&1
cat gpurun_out/r2l25_tune.txt
