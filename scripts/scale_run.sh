#!/bin/bash
# On a multi-GPU box (gpurun --gpus N): bench.py under torchrun at the given N for the given workloads.
#   N=8 bash scripts/scale_run.sh detection target
N=${N:-2}
mkdir -p gpurun_out
for wl in ${@:-detection}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $wl --steps 200 --warmup 20 > gpurun_out/scale_${wl}_n$N.json 2> gpurun_out/scale_${wl}_n$N.err
  tail -2 gpurun_out/scale_${wl}_n$N.err
  python - $wl $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/scale_%s_n%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print(sys.argv[1], "N", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 5), "e2e", round(d["e2e"]["value"]),
      "parity", d["parity_check"]["result"], "gather", d.get("gather_check"), d["config"].get("parallelism"))
PY
done
