#!/bin/bash
O=gpurun_out/r4m; mkdir -p $O
for i in 1 2 3; do
  for v in cur r32; do
    if [ $v = r32 ]; then export NUBOVCA_LIB=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_r32.so; else unset NUBOVCA_LIB; fi
    python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench_${v}_$i.json 2> $O/bench_${v}_$i.err
    python -c "
import json;d=json.load(open('$O/bench_${v}_$i.json'));print('$v',round(d['value'],1),round(d['e2e']['value'],1),round(d['stage_ms_isolated']['pyramid_rowscan'],4))"
  done
done
