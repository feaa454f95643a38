#!/usr/bin/env bash
# compute-sanitizer over the kernel-level GPU tests (memcheck: out-of-bounds / misaligned accesses; racecheck: shared
# memory hazards of the mbarrier / named-barrier protocols).  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
mkdir -p gpurun_out
SEL="tests/test_gpu_00_ops.py tests/test_gpu_01_conv_tc.py tests/test_gpu_02_elementwise.py"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool ($(date +%T))"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $SEL -m gpu -q -x -p no:cacheprovider \
      -k "not dgrad_mn_major and not gemm_mn" > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"
  grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard" gpurun_out/sanitize_$tool.log | tail -8
done
