#!/bin/bash
# round 2, lease 2: grouped + packed fused kernel and batched frame prep -- parity, variant sweep, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -40 > gpurun_out/r2l2_pytest.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=4
{
NICP_TILE_CONFIG=4 python tools/tune_corr.py
for g in 1 2 4 8 16; do for mb in 16 20; do
  echo "group=$g minb=$mb"; NICP_GROUP=$g NICP_GROUP_MINB=$mb python tools/tune_corr.py
done; done
} > gpurun_out/r2l2_tune.txt 2>&1
unset TUNE_CUR TUNE_CAND NICP_BATCH_SLOTS TUNE_REPS
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2l2_bench.json 2> gpurun_out/r2l2_bench.err
timeout 300 python tools/latency.py > gpurun_out/r2l2_latency.txt 2>&1
