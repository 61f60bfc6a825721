"""Per-call latency of the small cases (one stream, one frame in flight, page-locked host frames, graph replay):
config 1 (640x480 -> 160x120, element defaults), config 5's shape (1280x720 -> 640x360), a 96x64 ROI-sized detect and
config 2 (eyes inside faces, stand-in eye model).  Prints microseconds per call."""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402
import torch  # noqa: E402


def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def us(fn, n=400, warm=20):
    for _ in range(warm):
        fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return 1e6 * (time.perf_counter() - t) / n


src = os.path.join(ROOT, "nubomedia-vca_b200", "cascades")
casc = nv.Cascade(os.path.join(src, "haarcascade_frontalface_alt.xml"))
ctx = nv.Context(0, 1920, 1080)
f1 = pin(synth.frame(640, 480, 4, 1)); f5 = pin(synth.frame(1280, 720, 3, 1000))
roi = pin(synth.frame(96, 64, 1, 5, smin=0.5, smax=0.9)[..., 0])
out = {"cfg1_640x480_to_160": us(lambda: ctx.face_detect(casc, f1, 160, 1.25, 3, None)),
       "cfg5_1280x720_to_640": us(lambda: ctx.face_detect(casc, f5, 640, 1.25, 3, None)),
       "roi_96x64_detect": us(lambda: ctx.detect_multiscale(casc, roi, 1.1, 2, (20, 20)))}
d = tempfile.mkdtemp()
shutil.copy(os.path.join(src, "haarcascade_frontalface_alt.xml"), d)
for f in ("haarcascade_mcs_lefteye.xml", "haarcascade_mcs_righteye.xml"):
    shutil.copy(os.path.join(src, "haarcascade_eye.xml"), os.path.join(d, f))
f2 = pin(synth.frame(1280, 720, 3, 2, smin=0.4, smax=0.6))
e = nv.Element("nuboeyedetector", 0, d)
out["cfg2_eyes_1280x720"] = us(lambda: e.process(f2), 150)
e.close(); ctx.close(); shutil.rmtree(d, ignore_errors=True)
print({k: round(v, 1) for k, v in out.items()})
