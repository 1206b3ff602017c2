#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l46_pytest.log 2>&1
tail -2 gpurun_out/r2l46_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2l46_smoke.txt 2>&1
tail -1 gpurun_out/r2l46_smoke.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l46_bench_ref.json 2> gpurun_out/r2l46_bench_ref.err
timeout 1200 python bench.py > gpurun_out/r2l46_bench.json 2> gpurun_out/r2l46_bench.err
