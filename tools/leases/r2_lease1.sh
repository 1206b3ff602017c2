#!/bin/bash
# round 2, lease 1: parity on hardware (0 skipped), the reference's own CUDA iteration beside ours, baseline bench,
# ncu --set full of the fused kernel / frame-prep kernels / solve, FP32 pipe microbenchmark
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2l1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -25 > gpurun_out/r2l1_pytest.log
timeout 300 python tests/compare_ref_pwn_cuda.py > gpurun_out/r2l1_compare_ref_pwn_cuda.json 2> gpurun_out/r2l1_compare_ref_pwn_cuda.err
timeout 120 tools/microbench/fp32_pipes > gpurun_out/r2l1_fp32_pipes.txt 2>&1
timeout 600 python bench.py > gpurun_out/r2l1_bench.json 2> gpurun_out/r2l1_bench.err
timeout 300 python tools/bench_configs.py > gpurun_out/r2l1_configs.json 2> gpurun_out/r2l1_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corr_lin_tiled -s 12 -c 2 -o gpurun_out/r2l1_corr \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2l1_ncu_corr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stats|k_integral|k_depth_convert' -s 16 -c 8 \
  -o gpurun_out/r2l1_prep python tools/latency.py > gpurun_out/r2l1_ncu_prep.log 2>&1
NICP_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_reduce_solve|k_project|k_corr_lin' -s 9 -c 9 \
  -o gpurun_out/r2l1_single python tools/latency.py > gpurun_out/r2l1_ncu_single.log 2>&1
ls -la gpurun_out
