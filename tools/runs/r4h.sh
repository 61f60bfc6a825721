#!/bin/bash
O=gpurun_out/r4h; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_elements.py -m gpu -x -q > $O/gputests.log 2>&1; tail -2 $O/gputests.log
for i in 1 2 3; do python tools/small_frame_latency.py 2>&1 | tail -1; done
python tools/small_breakdown.py 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_small.csv python tools/ncu_small.py > $O/ncu_s.log 2>&1
