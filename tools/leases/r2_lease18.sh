#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l18_pytest.log 2>&1
tail -30 gpurun_out/r2l18_pytest.log
