#!/bin/bash
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_map_ops.py tests/test_host_cpp.py -q -m gpu -x \
  -k "not full_size and not 640 and not loop_closure and not real_kinect and not teacher_forced_full and not visible_gpus" > gpurun_out/r2l42_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2l42_memcheck.log
tail -4 gpurun_out/r2l42_memcheck.log
NICP_GROUP_MIN_AVG=0 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
  -k "grouped or priors_in_a_batch or determinism or batched_prep or random_scenes" > gpurun_out/r2l42_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2l42_racecheck.log
tail -4 gpurun_out/r2l42_racecheck.log
NICP_GROUP_MIN_AVG=0 timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
  -k "grouped or priors_in_a_batch or determinism or batched_prep" > gpurun_out/r2l42_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/r2l42_synccheck.log
tail -4 gpurun_out/r2l42_synccheck.log
