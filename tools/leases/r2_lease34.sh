#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l34_pytest.log 2>&1
tail -3 gpurun_out/r2l34_pytest.log
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
timeout 300 python tools/tune_corr.py | tail -1 > gpurun_out/r2l34_tune.txt 2>&1
cat gpurun_out/r2l34_tune.txt
timeout 900 python bench.py --no-cpu-baseline --no-configs > gpurun_out/r2l34_bench.json 2> gpurun_out/r2l34_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2l34_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['clocks'])"
