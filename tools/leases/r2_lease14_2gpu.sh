#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2l14_gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sharded" > gpurun_out/r2l14_pytest_sharded.log 2>&1
tail -3 gpurun_out/r2l14_pytest_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l14_bench_2gpu.json 2> gpurun_out/r2l14_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2l14_ref_2gpu.json 2> gpurun_out/r2l14_ref_2gpu.err
