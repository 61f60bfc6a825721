"""Median CUDA-event time of the fused tracker kernel on bench.py's 720p sequence (and a 1080p one): python tools/trk_time.py"""
import os
import sys

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
    ctx.close()
print(out)
