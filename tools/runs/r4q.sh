#!/bin/bash
O=gpurun_out/r4q; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; tail -2 $O/gputests.log
python tools/fuzz_parity.py 30 31 2>&1 | tail -1
for i in 1 2 3; do python tools/small_frame_latency.py 2>&1 | tail -1; done
python tools/small_breakdown.py 2>&1 | tail -1
for i in 1 2; do
python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err
python -c "
import json;d=json.load(open('$O/bench_$i.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['gpu_launches'])"
done
