#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r2l32_tracking.json 2> gpurun_out/r2l32_tracking.err <<'PY'
import json, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs
from g2o_frontend_b200 import capi
ctx = capi.Context(0)
print(json.dumps(bench_configs.config3(ctx, None, 1000)))
PY
cat gpurun_out/r2l32_tracking.json; tail -3 gpurun_out/r2l32_tracking.err
