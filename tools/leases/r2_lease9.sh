#!/bin/bash
set -x
mkdir -p gpurun_out
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=4
{
echo "warps=1 group=16"; NICP_GROUP=16 python tools/tune_corr.py
for mb in 16 20; do for g in 16 32; do
  echo "warps=2 minb=$mb group=$g"; NICP_GROUP_WARPS=2 NICP_GROUP_MINB=$mb NICP_GROUP=$g python tools/tune_corr.py
done; done
} > gpurun_out/r2l9_tune.txt 2>&1
unset TUNE_CUR TUNE_CAND NICP_BATCH_SLOTS TUNE_REPS
NICP_GROUP_WARPS=2 NICP_GROUP=32 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "determinism or correspondence_and or batch or inner or priors or epoch" 2>&1 | tail -5 > gpurun_out/r2l9_pytest_w2.log
