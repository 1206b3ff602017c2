#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "determinism or batch or grouped or priors or epoch or sharded or loop_closure" > gpurun_out/r2l24_pytest.log 2>&1
tail -3 gpurun_out/r2l24_pytest.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
for x in 0 1 2 3 4 6; do
  echo "NICP_FUSE_PROJECT=$x"; NICP_FUSE_PROJECT=$x timeout 300 python tools/tune_corr.py | tail -1
done > gpurun_out/r2l24_tune.txt 2>&1
cat gpurun_out/r2l24_tune.txt
