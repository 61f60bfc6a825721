#!/bin/bash
O=gpurun_out/r4l; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -m gpu -x -q > $O/gputests.log 2>&1; tail -2 $O/gputests.log
for i in 1 2 3; do
  for v in new prev; do
    if [ $v = prev ]; then export NUBOVCA_LIB=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_prev.so; else unset NUBOVCA_LIB; fi
    python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench_${v}_$i.json 2> $O/bench_${v}_$i.err
    python -c "
import json;d=json.load(open('$O/bench_${v}_$i.json'));print('$v',round(d['value'],1),round(d['e2e']['value'],1),{k:round(x,4) for k,x in d['stage_ms_isolated'].items() if k in ('cascade_stage0','cascade_tiles')})"
  done
done
unset NUBOVCA_LIB
python tools/small_frame_latency.py 2>&1 | tail -1
