"""A 96x64 ROI-sized detect and a config-1 frame with plain launches (ctx debug mode keeps the graph off): the command
for the launch lists of the small cases in profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402

casc = nv.Cascade(os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml"))
ctx = nv.Context(0, 1920, 1080)
ctx.set_profile(True)
roi = synth.frame(96, 64, 1, 5, smin=0.5, smax=0.9)[..., 0].copy()
f1 = synth.frame(640, 480, 4, 1)
for _ in range(3):
    a = ctx.detect_multiscale(casc, roi, 1.1, 2, (20, 20))
for _ in range(3):
    b = ctx.face_detect(casc, f1, 160, 1.25, 3, None)
print(len(a), len(b), ctx.stage_times())
ctx.close()
