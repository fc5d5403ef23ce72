#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): collects everything profiles/ is built from into gpurun_out/ (round tag $R).
#   one bench line per workload + the reference arm, the ncu launch list of the bench command, ncu --set full of the
#   hot kernels (source-level), per-kernel DRAM traffic.
R=${R:-r2}
set -x
mkdir -p gpurun_out
for wl in detection target ssd300 dspnet_cs nms detection_heads train_tail; do
  python bench.py --workload $wl > gpurun_out/bench_${wl}_$R.json 2> gpurun_out/bench_${wl}_$R.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_detection_$R.err
python bench.py --impl reference --workload target --steps 5 --warmup 1 > gpurun_out/bench_ref_target_$R.json 2>> gpurun_out/bench_target_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-soak --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_target_$R.csv \
    python bench.py --workload target --steps 5 --warmup 3 --no-cpu-baseline --no-soak --no-e2e > gpurun_out/ncu_bench_t.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"det_stream|det_sort|det_pair|target_stream|target_match" -c 10 \
    -o gpurun_out/prof_all_$R python scripts/prof_once.py both 2 > gpurun_out/ncu_all.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"det_stream_heads" -c 1 \
    -o gpurun_out/prof_heads_$R python scripts/prof_heads.py > gpurun_out/ncu_heads.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nms_cull|nms_resolve" -c 6 --launch-skip 200 \
    -o gpurun_out/prof_nms_$R python scripts/nms_breakdown.py 200000 > gpurun_out/ncu_nms.log 2>&1
tail -2 gpurun_out/ncu_all.log gpurun_out/ncu_heads.log gpurun_out/ncu_nms.log
