#!/bin/bash
# A/B on the GPU box: target matcher shortlist (knob 14) and detection class-tile staging (knob 12).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "target or staging" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_quoted_configs.py tests/test_loss.py -m gpu -x -q -k "target or loss or tail" 2>&1 | tail -5
show() {
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
print(sys.argv[1], "VALUE", round(d["value"]), "us/step", round(d["ms_per_step"] * 1000, 2), "kernels",
      {k: round(v * 1000, 1) for k, v in (r.get("all_kernels_ms") or {}).items()}, "parity", (d.get("parity_check") or {}).get("result"))
PY
}
for v in 0 1; do
  TUNE14=$v timeout 300 python scripts/bench_knob.py --workload target --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ab_short_$v.json 2>gpurun_out/ab_short_$v.err
  show gpurun_out/ab_short_$v.json
done
for v in 1 2; do
  TUNE12=$v timeout 300 python scripts/bench_knob.py --workload detection --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ab_lean_$v.json 2>gpurun_out/ab_lean_$v.err
  show gpurun_out/ab_lean_$v.json
done
