#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l16_pytest.log 2>&1
tail -4 gpurun_out/r2l16_pytest.log
timeout 300 python tools/latency.py | tail -1 > gpurun_out/r2l16_latency.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline --no-configs > gpurun_out/r2l16_bench.json 2> gpurun_out/r2l16_bench.err
