#!/bin/bash
python -m pytest tests -m gpu -x -q -k "track or trk" 2>&1 | tail -2
python tools/fuzz_tracker.py 20 7 2>&1 | tail -1
python tools/fuzz_ref_elements.py 2>&1 | tail -1
for i in 1 2; do for m in 1 0; do echo "HOST_WRITES=$m $(NUBOVCA_HOST_WRITES=$m python tools/trk_time.py 2>&1 | tail -1)"; done; done
