#!/bin/bash
O=gpurun_out/r4j; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "track or trk" > $O/gputests.log 2>&1; tail -2 $O/gputests.log
python tools/fuzz_tracker.py 30 3 2>&1 | tail -2
for i in 1 2; do
echo "new $(python tools/trk_time.py 2>&1 | tail -1)"
echo "old $(NUBOVCA_LIB=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_old.so python tools/trk_time.py 2>&1 | tail -1)"
done
