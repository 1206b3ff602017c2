#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2l44_bench.json 2> gpurun_out/r2l44_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l44_bench_ref.json 2> gpurun_out/r2l44_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/r2l44_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --currents 8 --candidates 128 > gpurun_out/r2l44_ncu_launches.log 2>&1
NICP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 80 --csv --log-file gpurun_out/r2l44_single_launches.csv python tools/latency.py > /dev/null 2>&1
timeout 300 python tools/latency.py | tail -1 > gpurun_out/r2l44_latency.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2l44_smoke.txt 2>&1
tail -1 gpurun_out/r2l44_smoke.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_corr_lin_group -s 22 -c 2 -o gpurun_out/r2l44_corr_bench \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --currents 8 --candidates 128 > gpurun_out/r2l44_ncu_corr.log 2>&1
