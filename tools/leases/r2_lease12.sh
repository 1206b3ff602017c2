#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_corr_lin_group -s 22 -c 2 -o gpurun_out/r2l12_corr_bench \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --currents 8 --candidates 128 > gpurun_out/r2l12_ncu_corr.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/r2l12_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --currents 8 --candidates 128 > gpurun_out/r2l12_ncu_launches.log 2>&1
PREP_FRAMES=16 PREP_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stats|k_integral|k_depth_convert' -s 24 -c 4 \
  -o gpurun_out/r2l12_prep_batch python tools/prep_batch_timing.py > gpurun_out/r2l12_ncu_prep.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -q -k "stage" 2>&1 | tail -3 > gpurun_out/r2l12_pytest.log
