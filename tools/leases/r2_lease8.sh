#!/bin/bash
set -x
mkdir -p gpurun_out
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=4
{
for v in 1 2 3 4; do for g in 16 8; do
  echo "variant=$v group=$g"; NICP_CORR_VARIANT=$v NICP_GROUP=$g python tools/tune_corr.py
done; done
echo "variant=1 group=1"; NICP_CORR_VARIANT=1 NICP_GROUP=1 python tools/tune_corr.py
echo "variant=3 group=1"; NICP_CORR_VARIANT=3 NICP_GROUP=1 python tools/tune_corr.py
} > gpurun_out/r2l8_tune.txt 2>&1
unset TUNE_CUR TUNE_CAND NICP_BATCH_SLOTS TUNE_REPS
for v in 1 3; do
NICP_CORR_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "determinism or correspondence_and or batch or inner or priors or epoch" 2>&1 | tail -5 > gpurun_out/r2l8_pytest_v$v.log
done
