#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l11_pytest.log 2>&1
tail -5 gpurun_out/r2l11_pytest.log
timeout 300 python tools/prep_batch_timing.py > gpurun_out/r2l11_prep.json 2> gpurun_out/r2l11_prep.err
NICP_PREP_BATCH=4 timeout 300 python tools/prep_batch_timing.py > gpurun_out/r2l11_prep_b4.json 2>> gpurun_out/r2l11_prep.err
NICP_PREP_STREAM_FROM=100 timeout 300 python tools/prep_batch_timing.py > gpurun_out/r2l11_prep_smemcols.json 2>> gpurun_out/r2l11_prep.err
timeout 300 python tools/latency.py > gpurun_out/r2l11_latency.txt 2>&1
PREP_FRAMES=16 PREP_REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2l11_prep_launches.csv python tools/prep_batch_timing.py > /dev/null 2>&1
NICP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 80 --csv --log-file gpurun_out/r2l11_single_launches.csv python tools/latency.py > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/r2l11_bench.json 2> gpurun_out/r2l11_bench.err
