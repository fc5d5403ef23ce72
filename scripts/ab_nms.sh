#!/bin/bash
# On the GPU box: NMS tests, then the nms workload with the chained launches on (default) and off (knob 17).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_quoted_configs.py -m gpu -x -q -k "nms" 2>&1 | tail -2
for v in ${VARIANTS:-1 0}; do
  TUNE17=$v timeout 300 python scripts/bench_knob.py --workload nms --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/ab_nms_$v.json 2>gpurun_out/ab_nms_$v.err
  python - $v <<'PY'
import json, sys
d = json.loads(open("gpurun_out/ab_nms_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("chained", sys.argv[1], "ms/sweep", round(d["ms_per_step"], 3), [(r["n"], round(r["force_ms"], 3), round(r["per_class_ms"], 3)) for r in d["sweep"]], (d.get("parity_check") or {}).get("result"))
PY
done
