#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l35_pytest.log 2>&1
tail -3 gpurun_out/r2l35_pytest.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
timeout 300 python tools/tune_corr.py | tail -1 > gpurun_out/r2l35_tune.txt 2>&1
cat gpurun_out/r2l35_tune.txt
timeout 300 python tools/latency.py | tail -1 > gpurun_out/r2l35_latency.txt 2>&1
cat gpurun_out/r2l35_latency.txt
