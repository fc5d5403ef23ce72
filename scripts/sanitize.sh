#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU parity tests (run under gpurun).
SEL=${SEL:-'test_detection_presets and ssd300 or test_target_presets and ssd300 or test_detection_dense_and_ties or test_target_adversarial_matching or test_graph_cache_replays_are_identical'}
for tool in ${TOOLS:-memcheck racecheck initcheck}; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/san_$tool.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|Race reported|Uninitialized" gpurun_out/san_$tool.log | sort | uniq -c | sort -rn | head -8
  grep -E -A12 "Uninitialized|Race reported|Invalid" gpurun_out/san_$tool.log | grep -E "Uninitialized|Race|Invalid|at .*\.cu|in .*kernel|Saved host|by thread|Access" | head -${DETAIL:-24}
done
