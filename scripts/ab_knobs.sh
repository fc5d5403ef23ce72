#!/bin/bash
# On the GPU box: bench.py --workload $WL under knob settings given as arguments, each "k=v[,k=v...]" ("-" = defaults).
mkdir -p gpurun_out
WL=${WL:-detection}
for spec in "$@"; do
  envs=""
  if [ "$spec" != "-" ]; then for kv in ${spec//,/ }; do envs="$envs TUNE${kv%%=*}=${kv##*=}"; done; fi
  env $envs timeout 300 python scripts/bench_knob.py --workload $WL --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ab_k.json 2>gpurun_out/ab_k.err
  python - "$spec" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_k.json"))
r = d.get("roofline", {})
print("knobs", sys.argv[1], "us/step", round(d["ms_per_step"] * 1000, 2), "kernels",
      {k: round(v * 1000, 1) for k, v in (r.get("all_kernels_ms") or {}).items()}, "parity", (d.get("parity_check") or {}).get("result"))
PY
done
