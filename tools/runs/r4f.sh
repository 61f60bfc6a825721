#!/bin/bash
O=gpurun_out/r4f; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; tail -3 $O/gputests.log
python tools/fuzz_parity.py 30 21 2>&1 | tail -3
python tools/fuzz_elements.py 2>&1 | tail -2
for i in 1 2; do python tools/small_frame_latency.py 2>&1 | tail -1; done
python tools/small_breakdown.py 2>&1 | tail -1
python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.load(open('$O/bench.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['gpu_launches'],{k:round(x,4) for k,x in d['stage_ms_isolated'].items()})"
