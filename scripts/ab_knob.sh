#!/bin/bash
# A/B of the detection stream-kernel variants on the bench workload: TUNE0 values given as arguments.
for v in "$@"; do
  TUNE0=$v timeout 300 python scripts/bench_knob.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_$v.json'));print('VARIANT',$v,round(d['value']),d['ms_per_step'],d['roofline']['all_kernels_ms'],d['roofline']['frac'])"
done
