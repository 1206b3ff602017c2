#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l10_pytest.log 2>&1
tail -5 gpurun_out/r2l10_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2l10_bench.json 2> gpurun_out/r2l10_bench.err
timeout 300 python tools/latency.py > gpurun_out/r2l10_latency.txt 2>&1
export TUNE_CUR=64 TUNE_CAND=4 NICP_BATCH_SLOTS=256 TUNE_REPS=3
{ echo "no sharing shape: 64 currents x 4 candidates (per-pair kernel)"; python tools/tune_corr.py
  echo "forced grouped"; NICP_GROUP_MIN_AVG=0 python tools/tune_corr.py; } > gpurun_out/r2l10_tune_nosharing.txt 2>&1
