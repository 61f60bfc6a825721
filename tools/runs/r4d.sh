#!/bin/bash
O=gpurun_out/r4d; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -m gpu -x -q > $O/gputests.log 2>&1; tail -3 $O/gputests.log
for m in 0 2 4; do echo "TAIL_TAB=$m"; NUBOVCA_TAIL_TAB=$m python tools/small_frame_latency.py 2>&1 | tail -1; done
for i in 1 2; do for m in 0 1; do
NUBOVCA_TAIL_TAB=$m python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench_tt$m.json 2> $O/bench_tt$m.err
python -c "
import json;d=json.load(open('$O/bench_tt$m.json'));print('tail_tab=$m',round(d['value'],1),round(d['e2e']['value'],1),{k:round(x,4) for k,x in d['stage_ms_isolated'].items()})"
done; done
