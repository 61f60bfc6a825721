#!/bin/bash
# the driver's round-end sequence: reference arm, own arm (default flags), smoke
O=gpurun_out/r4i; mkdir -p $O
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -c "
import json;d=json.load(open('$O/bench.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['gpu_launches'],d['parity_in_run'],d['roofline']['frac'],d['cpu_baseline'])"
