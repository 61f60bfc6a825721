"""One cfg5 frame handed over as NV12 planes (1280x720 -> 640x360), a few repetitions on one context: the command the
ncu captures of the ingest kernel in profiles/ were taken with (run it without ncu first)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402

casc = nv.Cascade(os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml"))
ctx = nv.Context(0, 1920, 1080)
ctx.set_profile(True)            # plain launches (no graph), so that ncu sees every kernel by name
for (w, h, w2p) in ((1280, 720, 640), (1920, 1080, 1920)):
    buf = synth.to_yuv420(synth.frame(w, h, 3, 1000), "NV12")
    planes = synth.yuv420_planes(buf, w, h, "NV12")
    for _ in range(3):
        r = ctx.face_detect_yuv(casc, planes, "NV12", w2p, 1.25, 3, None)
    print(w, h, len(r), ctx.stage_times())
ctx.close()
