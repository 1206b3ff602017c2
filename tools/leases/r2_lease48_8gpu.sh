#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2l48_gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sharded" > gpurun_out/r2l48_pytest_sharded.log 2>&1
tail -3 gpurun_out/r2l48_pytest_sharded.log
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-configs > gpurun_out/r2l48_bench_${n}gpu.json 2> gpurun_out/r2l48_bench_${n}gpu.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2l48_ref_8gpu.json 2> gpurun_out/r2l48_ref_8gpu.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2l48_bench_1gpu.json 2> gpurun_out/r2l48_bench_1gpu.err
