#!/bin/bash
# A/B on the GPU box: late L2 prefetch distance of the target stream kernel (knob 10).
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
print(sys.argv[1], "VALUE", round(d["value"]), "us/step", round(d["ms_per_step"] * 1000, 2), "kernels",
      {k: round(v * 1000, 1) for k, v in (r.get("all_kernels_ms") or {}).items()}, "parity", (d.get("parity_check") or {}).get("result"))
PY
}
for v in "$@"; do
  TUNE10=$v timeout 300 python scripts/bench_knob.py --workload target --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ab_tpf_$v.json 2>gpurun_out/ab_tpf_$v.err
  show gpurun_out/ab_tpf_$v.json
done
