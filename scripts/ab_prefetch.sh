for d in 0 296 592 1036 1554 2072; do echo "prefetch $d"; python scripts/det_step_time.py 9=$d 2>&1 | tail -1; done
