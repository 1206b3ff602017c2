#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "teacher_forced or determinism" 2>&1 | tail -60 > gpurun_out/r2l3_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "end_to_end" 2>&1 | tail -40 > gpurun_out/r2l3_pytest_b.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=4
{
for g in 1 4 8 16; do for mb in 16 20; do
  echo "group=$g minb=$mb"; NICP_GROUP=$g NICP_GROUP_MINB=$mb python tools/tune_corr.py
done; done
} > gpurun_out/r2l3_tune.txt 2>&1
