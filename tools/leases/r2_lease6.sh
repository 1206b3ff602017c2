#!/bin/bash
set -x
mkdir -p gpurun_out
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=1
NICP_GROUP=16 NICP_GROUP_MINB=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corr_lin_group -s 12 -c 1 \
  -o gpurun_out/r2l6_group python tools/tune_corr.py > gpurun_out/r2l6_ncu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "end_to_end" > gpurun_out/r2l6_pytest.log 2>&1
