#!/bin/bash
mkdir -p gpurun_out
export TUNE_CUR=8 TUNE_CAND=64 TUNE_REPS=3
{
PROBE_CONTEXTS=1 NICP_BATCH_SLOTS=256 timeout 300 python tools/overlap_probe.py | tail -1
PROBE_CONTEXTS=2 NICP_BATCH_SLOTS=256 timeout 300 python tools/overlap_probe.py | tail -1
PROBE_CONTEXTS=2 NICP_BATCH_SLOTS=128 timeout 300 python tools/overlap_probe.py | tail -1
PROBE_CONTEXTS=4 NICP_BATCH_SLOTS=128 timeout 300 python tools/overlap_probe.py | tail -1
PROBE_CONTEXTS=1 NICP_BATCH_SLOTS=128 timeout 300 python tools/overlap_probe.py | tail -1
} > gpurun_out/r2l23_overlap.txt 2>&1
cat gpurun_out/r2l23_overlap.txt
