#!/bin/bash
mkdir -p gpurun_out
export TUNE_CUR=4 TUNE_CAND=64 NICP_BATCH_SLOTS=256 TUNE_REPS=5
for v in 0 1 0 1; do
  echo "NICP_PROJECT_BY_REFERENCE=$v"; NICP_PROJECT_BY_REFERENCE=$v timeout 300 python tools/tune_corr.py | tail -1
done > gpurun_out/r2l38_tune.txt 2>&1
cat gpurun_out/r2l38_tune.txt
timeout 900 python bench.py --no-cpu-baseline --no-configs > gpurun_out/r2l38_bench.json 2> gpurun_out/r2l38_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2l38_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['project_avg_launch_ms'],d['clocks'])"
