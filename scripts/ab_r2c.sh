#!/bin/bash
# On the GPU box: target tests, timeline, A/B of the PDL knob (15), all GPU tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "target or staging" 2>&1 | tail -5
python scripts/tgt_timeline.py 64 > gpurun_out/t64c.txt 2>&1; head -18 gpurun_out/t64c.txt
python scripts/tgt_timeline.py 8 > gpurun_out/t8c.txt 2>&1; head -2 gpurun_out/t8c.txt
python scripts/tgt_timeline.py 64 15=0 2>&1 | head -2
python scripts/tgt_timeline.py 8 15=0 2>&1 | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
