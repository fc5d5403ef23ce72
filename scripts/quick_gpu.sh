#!/bin/bash
# On the GPU box: parity tests + one short bench line per workload given as arguments (default: detection).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for wl in ${@:-detection}; do
  timeout 600 python bench.py --workload $wl --steps 100 --warmup 10 > gpurun_out/b_$wl.json 2>gpurun_out/b_$wl.err
  tail -3 gpurun_out/b_$wl.err
  python - $wl <<'PY'
import json, sys
d = json.load(open("gpurun_out/b_%s.json" % sys.argv[1]))
r = d.get("roofline", {})
print(sys.argv[1], "VALUE", round(d["value"]), "ms/step", d["ms_per_step"], "kernels", r.get("all_kernels_ms"), "frac", r.get("frac"),
      "whole", r.get("whole_op_frac"), "e2e", (d.get("e2e") or {}).get("value"), "parity", d.get("parity_check"),
      "cpu", d.get("cpu_baseline"), "launches/step", d.get("launches_per_step"))
PY
done
