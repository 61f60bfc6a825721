#!/bin/bash
# k_cascade_tail_tab on LARGE plans of moderate size (config 5's 640x360: 87 626 windows): latency and 32-stream throughput
O=gpurun_out/r4t; mkdir -p $O
for i in 1 2; do for m in 4 5; do
echo "TAIL_TAB=$m $(NUBOVCA_TAIL_TAB=$m python tools/small_frame_latency.py 2>&1 | tail -1)"
NUBOVCA_TAIL_TAB=$m python bench.py --steps 20 --no-cpu-baseline > $O/bench_$m.json 2> $O/bench_$m.err
python -c "
import json;d=json.load(open('$O/bench_$m.json'));a=d['aux'];print('  cfg3',round(d['value'],1),'cfg5 python',round(a['frames_per_s']),'nv12',round(a['nv12_ingest']['frames_per_s']),'native',round(a['native_host_threads']['bgr']['frames_per_s']),round(a['native_host_threads']['nv12']['frames_per_s']),'resident',round(a['device_resident']['frames_per_s']))"
done; done
