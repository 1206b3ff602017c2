#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stage_level" > gpurun_out/r2l13_pytest.log 2>&1
