"""Randomised parity run of the nubotracker path (not part of the test suite): random frame sizes, strides, thresholds,
area / distance parameters, noise levels and timestamp gaps (history decay) over random sequences, through long-lived
contexts, against the CPU oracle.  Usage: python tools/fuzz_tracker.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "nubomedia-vca_b200", "python"), os.path.join(ROOT, "oracle")]
import nubovca as nv  # noqa: E402
import oracle as O  # noqa: E402
from nubovca import synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
t0 = time.time()
nseq = nframes = bad = 0
while time.time() - t0 < budget:
    W = int(rng.integers(140, 700)) * 2; H = int(rng.integers(140, 380)) * 2
    n = int(rng.integers(3, 9))
    frames = synth.tracker_sequence(W, H, n, seed=int(rng.integers(1 << 30)), nsq=int(rng.integers(1, 5)),
                                    noise=int(rng.choice([0, 0, 20, 200, 2000])))
    thr = int(rng.choice([5, 20, 20, 60])); mina = int(rng.choice([0, 50, 400])); maxa = int(rng.choice([3000, 30000, 10 ** 7]))
    dist = int(rng.choice([0, 35, 120]))
    ctx = nv.Context(0, W, H)
    st = O.TrackerState(W, H)
    ts = 0.0
    for f in frames:
        ts += float(rng.choice([33.3, 33.3, 100.0, 450.0]))            # gaps beyond the 200 ms history window too
        a = ctx.tracker_process(f, ts, thr, mina, maxa, dist)
        b, _, _ = st.process(f, ts, thr, mina, maxa, dist)
        nframes += 1
        if a.shape != b.shape or not (a == b).all():
            bad += 1
            print("MISMATCH", W, H, thr, mina, maxa, dist, ts, a.tolist()[:4], b.tolist()[:4], flush=True)
    ctx.close()
    nseq += 1
print(f"fuzz_tracker: {nseq} sequences, {nframes} frames, {bad} mismatches, seed {seed}, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
