#!/bin/bash
O=gpurun_out/r4p; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_elements.py tests/test_fuzz_gpu.py tests/test_yuv_ingest.py -m gpu -x -q > $O/gputests.log 2>&1; tail -2 $O/gputests.log
for i in 1 2 3; do for m in 1 0; do echo "HOST_WRITES=$m $(NUBOVCA_HOST_WRITES=$m python tools/small_frame_latency.py 2>&1 | tail -1)"; done; done
for i in 1 2; do for m in 1 0; do
NUBOVCA_HOST_WRITES=$m python bench.py --steps 60 --no-aux --no-cpu-baseline > $O/bench_$m.json 2> $O/bench_$m.err
python -c "
import json;d=json.load(open('$O/bench_$m.json'));print('host_writes=$m',round(d['value'],1),round(d['e2e']['value'],1))"
done; done
