#!/bin/bash
L=$PWD/nubomedia-vca_b200/lib/ab/libnubovca_trk16.so
NUBOVCA_LIB=$L python -m pytest tests -m gpu -x -q -k "track or trk" 2>&1 | tail -2
NUBOVCA_LIB=$L python tools/fuzz_tracker.py 20 5 2>&1 | tail -1
for i in 1 2; do
echo "th32 $(python tools/trk_time.py 2>&1 | tail -1)"
echo "th16 $(NUBOVCA_LIB=$L python tools/trk_time.py 2>&1 | tail -1)"
done
