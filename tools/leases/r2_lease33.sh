#!/bin/bash
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_map_ops.py tests/test_host_cpp.py -q -m gpu -x \
  -k "not full_size and not 640 and not loop_closure and not real_kinect and not teacher_forced_full and not visible_gpus" > gpurun_out/r2l33_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2l33_memcheck.log
tail -6 gpurun_out/r2l33_memcheck.log
NICP_GROUP_MIN_AVG=0 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
  -k "grouped or priors_in_a_batch or determinism or batched_prep" > gpurun_out/r2l33_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2l33_racecheck.log
tail -6 gpurun_out/r2l33_racecheck.log
