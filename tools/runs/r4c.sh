#!/bin/bash
O=gpurun_out/r4c; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; tail -3 $O/gputests.log
for m in 1 2; do echo PDL=$m; NUBOVCA_PDL=$m python tools/fuzz_parity.py 25 11 2>&1 | tail -3; done
for m in 0 1; do echo "PDL=$m"; NUBOVCA_PDL=$m python tools/small_frame_latency.py 2>&1 | tail -1; done
for m in 1 2; do
NUBOVCA_PDL=$m python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench_pdl$m.json 2> $O/bench_pdl$m.err
python -c "
import json;d=json.load(open('$O/bench_pdl$m.json'));print('pdl$m',round(d['value'],1),round(d['e2e']['value'],1),{k:round(x,4) for k,x in d['stage_ms_isolated'].items()})"
done
