#!/bin/bash
# A/B of the round-4 changes (4-column row scan, PDL on small plans) against the previous library
O=gpurun_out/r4b; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; tail -3 $O/gputests.log
for m in 0 1; do echo "PDL=$m"; NUBOVCA_PDL=$m python tools/small_frame_latency.py 2>&1 | tail -1; done
echo old; NUBOVCA_LIB=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_old.so python tools/small_frame_latency.py 2>&1 | tail -1
for i in 1 2; do
  for v in new old; do
    if [ $v = old ]; then export NUBOVCA_LIB=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_old.so; else unset NUBOVCA_LIB; fi
    python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench_${v}_$i.json 2> $O/bench_${v}_$i.err
    python -c "
import json;d=json.load(open('$O/bench_${v}_$i.json'));print('$v',round(d['value'],1),round(d['e2e']['value'],1),{k:round(x,4) for k,x in d['stage_ms_isolated'].items()})"
  done
done
