#!/bin/bash
O=gpurun_out/r4k; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_elements.py tests/test_fuzz_gpu.py -m gpu -x -q > $O/gputests.log 2>&1; tail -2 $O/gputests.log
for i in 1 2 3; do for m in 4 12; do echo "TAIL_TAB=$m $(NUBOVCA_TAIL_TAB=$m python tools/small_frame_latency.py 2>&1 | tail -1)"; done; done
