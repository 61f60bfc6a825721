"""Median CUDA-event time of the fused tracker kernel on bench.py's sequence at three frame sizes, and the wall time of a blocking
nv_tracker_process call on page-locked frames: python tools/trk_time.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402

out = {}
for w, h in ((1280, 720), (1920, 1080), (640, 360)):
    seq = synth.tracker_sequence(w, h, 8, seed=4)
    ctx = nv.Context(0, 1920, 1080)
    ctx.set_profile(True)
    k = []
    for i in range(300):
        ctx.tracker_process(seq[i % len(seq)], 33.3 * (i + 1))
        if i >= 20:
            k.append(ctx.tracker_kernel_ms())
    out[f"{w}x{h}"] = round(float(np.median(k)) * 1e3, 1)
    ctx.set_profile(False)
    import torch
    pseq = [torch.from_numpy(f).pin_memory().numpy() for f in seq]
    for i in range(50):
        ctx.tracker_process(pseq[i % len(pseq)], 33.3 * (400 + i))
    t = time.perf_counter()
    for i in range(400):
        ctx.tracker_process(pseq[i % len(pseq)], 33.3 * (500 + i))
    out[f"{w}x{h}_call_us"] = round(1e6 * (time.perf_counter() - t) / 400, 1)
    ctx.close()
print(out)
