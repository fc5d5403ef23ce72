#!/bin/bash
# compute-sanitizer over the GPU parity suite (run under gpurun).  torch's caching allocator is switched off so that an
# out-of-bounds access cannot hide inside its arena.  TOOLS / SEL select tools and tests; summaries land in gpurun_out/.
SEL=${SEL:-'not slow'}
for tool in ${TOOLS:-memcheck racecheck initcheck}; do
  echo "== $tool"
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout ${LIMIT:-1500} compute-sanitizer --tool $tool --error-exitcode 9 \
      python -m pytest tests -m gpu -q -k "$SEL" -p no:cacheprovider > gpurun_out/san_$tool.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san_$tool.log | tail -3
  grep -E "Invalid|Uninitialized|Race reported" gpurun_out/san_$tool.log | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//' | cut -c1-220 | sort | uniq -c | sort -rn | head -${DETAIL:-12}
done
