#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 1 2 1 2; do
  NICP_CORR_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline --no-configs --steps 8 --warmup 3 > gpurun_out/r2l45_bench_v$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r2l45_bench_v$v.json'));print('variant $v', round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), round(d['roofline']['avg_launch_ms'],4), d['clocks']['sm_mhz'])"
done > gpurun_out/r2l45_sustained.txt 2>&1
cat gpurun_out/r2l45_sustained.txt
