#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l15_pytest.log 2>&1
tail -4 gpurun_out/r2l15_pytest.log
for ft in 0 1; do for sw in 0 1; do
  echo "fuse_tail=$ft single_wave=$sw"; NICP_FUSE_TAIL=$ft NICP_SINGLE_WAVE=$sw timeout 300 python tools/latency.py | tail -1
done; done > gpurun_out/r2l15_latency.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l15_smoke.txt 2>&1
