#!/bin/bash
O=gpurun_out/r4e; mkdir -p $O
for i in 1 2 3; do for m in 0 4; do echo "TAIL_TAB=$m $(NUBOVCA_TAIL_TAB=$m python tools/small_frame_latency.py 2>&1 | tail -1)"; done; done
for v in tt0 tt8 tt16 tt32; do
  unset NUBOVCA_LIB; export NUBOVCA_TAIL_TAB=1
  [ $v = tt0 ] && export NUBOVCA_TAIL_TAB=0
  [ $v = tt16 ] && export NUBOVCA_LIB=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_tt16.so
  [ $v = tt32 ] && export NUBOVCA_LIB=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_tt32.so
  python bench.py --steps 40 --no-aux --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json;d=json.load(open('$O/bench_$v.json'));print('$v',round(d['value'],1),round(d['e2e']['value'],1),d['stage_ms_isolated']['cascade_tail'])"
done
