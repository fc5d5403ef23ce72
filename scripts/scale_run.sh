#!/bin/bash
# Runs ON THE GPU BOX under `gpurun --gpus 8`: the driver's scaling sequence N = 1, 2, 4, 8.
mkdir -p gpurun_out
for n in ${SCALE_NS:-1 2 4 8}; do
  if [ "$n" = "1" ]; then
    timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 200 --warmup 20 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1])
    print("SCALE", d["n_gpus"], round(d["value"]), d["ms_per_step"], d.get("gather_check"), round(d["e2e"]["value"]))
except Exception as e:
    print("SCALE n=$n failed", repr(e)[:200])
PY
done
