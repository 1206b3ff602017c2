#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "determinism or correspondence_and or batch or inner or priors or epoch" 2>&1 | tail -30 > gpurun_out/r2l7_pytest.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=4
{
for g in 1 4 16; do for mb in 16 20; do
  echo "group=$g minb=$mb"; NICP_GROUP=$g NICP_GROUP_MINB=$mb python tools/tune_corr.py
done; done
} > gpurun_out/r2l7_tune.txt 2>&1
export TUNE_REPS=1
NICP_GROUP=16 NICP_GROUP_MINB=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corr_lin_group -s 12 -c 1 \
  -o gpurun_out/r2l7_group python tools/tune_corr.py > gpurun_out/r2l7_ncu.log 2>&1
