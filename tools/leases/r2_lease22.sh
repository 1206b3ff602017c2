#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l22_pytest.log 2>&1
tail -4 gpurun_out/r2l22_pytest.log
timeout 900 python bench.py --no-configs > gpurun_out/r2l22_bench.json 2> gpurun_out/r2l22_bench.err
timeout 300 python tools/latency.py | tail -1 > gpurun_out/r2l22_latency.txt 2>&1
export TUNE_CUR=64 TUNE_CAND=4 NICP_BATCH_SLOTS=256 TUNE_REPS=3
{ echo "no sharing shape: 64 currents x 4 candidates (per-pair kernel)"; python tools/tune_corr.py | tail -1
  echo "forced grouped"; NICP_GROUP_MIN_AVG=0 python tools/tune_corr.py | tail -1; } > gpurun_out/r2l22_tune_nosharing.txt 2>&1
