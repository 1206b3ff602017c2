#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2l31_smoke.txt 2>&1
tail -2 gpurun_out/r2l31_smoke.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-configs > gpurun_out/r2l31_bench_2gpu.json 2> gpurun_out/r2l31_bench_2gpu.err
