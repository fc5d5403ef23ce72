#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): collects everything profiles/ is built from into gpurun_out/.
#   bench line + reference arm, ncu launch list of the bench command, ncu --set full of the stream kernels and of
#   the latency-bound kernels (source-level), per-kernel dram traffic.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1.json 2>> gpurun_out/bench_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-soak > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"det_stream|det_sort|det_nms|target_stream|target_match" -c 10 \
    -o gpurun_out/prof_all_r1 python scripts/prof_once.py both 2 > gpurun_out/ncu_all.log 2>&1
tail -2 gpurun_out/ncu_all.log
