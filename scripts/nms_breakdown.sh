#!/bin/bash
# NMS kernel duration (ncu, cold) with the kernel cut short at successive points (DSPMB_NMS_DEBUG: 4 launch only,
# 2 after the member lists, 1 lists + staging + resolve without pair tests, 0 full).
for d in 4 2 1 0; do
  DSPMB_NMS_DEBUG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:det_nms --csv \
    --log-file gpurun_out/nb_$d.csv python scripts/prof_once.py det 3 > /dev/null 2>&1
  echo "DEBUG $d $(grep det_nms gpurun_out/nb_$d.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
