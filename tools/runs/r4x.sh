#!/bin/bash
O=gpurun_out/r4x; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_elements.py tests/test_fuzz_gpu.py -m gpu -x -q > $O/gputests.log 2>&1; tail -2 $O/gputests.log
python tools/fuzz_parity.py 30 61 2>&1 | tail -1
for i in 1 2 3; do python tools/small_frame_latency.py 2>&1 | tail -1; done
python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.load(open('$O/bench.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['stage_ms_isolated']['group_rectangles'])"
python tools/general_cascade_latency.py 2>&1 | tail -3
