#!/bin/bash
# long fuzz runs on the round's final library
python tools/fuzz_parity.py 150 101 2>&1 | tail -2
python tools/fuzz_tracker.py 60 102 2>&1 | tail -1
python tools/fuzz_elements.py 2>&1 | tail -1
python tools/fuzz_ref_elements.py 2>&1 | tail -2
