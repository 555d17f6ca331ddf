#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (SURVEY.md §5 "race detection / sanitizers" row).
# usage (under gpurun): bash profiles/probes/r2_sanitize.sh   -> gpurun_out/r3_memcheck.log, r3_racecheck.log, r3_synccheck.log
SEL='exclusive_and_gang or scarce_capacity or degenerate_fleets or exclusive_group_with_nothing or small_deltas or plan_dense_matrix_and_lists or ingested_snapshot'
FILES="tests/test_gpu_parity.py tests/test_gpu_groups.py tests/test_gpu_delta.py tests/test_gpu_callers.py"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 1 \
      python -m pytest $FILES -m gpu -x -q -k "$SEL" > gpurun_out/r3_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r3_$tool.log | tail -3
done
