#!/bin/bash
# On the GPU box: parity tests + one short bench line, summary on stdout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err
tail -3 gpurun_out/b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/b.json"))
print("VALUE", round(d["value"]), d["ms_per_step"], d["roofline"]["all_kernels_ms"], d["roofline"]["frac"], d["e2e"]["value"],
      d.get("target", {}).get("images_per_s"), d.get("target", {}).get("kernel_ms"))
PY
