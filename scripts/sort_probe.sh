#!/bin/bash
for d in 0 8; do
  DSPMB_NMS_DEBUG=$d timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-soak > gpurun_out/sp_$d.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sp_$d.json'));print('DEBUG',$d,d['ms_per_step'],d['roofline']['all_kernels_ms'])"
done
