#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l19_pytest.log 2>&1
tail -6 gpurun_out/r2l19_pytest.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
timeout 300 python tools/tune_corr.py > gpurun_out/r2l19_tune.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline --no-configs > gpurun_out/r2l19_bench.json 2> gpurun_out/r2l19_bench.err
