"""Three BASELINE config-3 frames (1920x1080 BGR, processing width 1920, sf 1.1, min 24x24) on one context with plain
launches: the command the `--set full` captures of the non-cascade kernels in profiles/ were taken with (run it without
ncu first)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402

casc = nv.Cascade(os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml"))
ctx = nv.Context(0, 1920, 1080)
ctx.set_profile(True)            # plain launches (no graph), so that ncu sees every kernel by name
fr = synth.frame(1920, 1080, 6, 3)
for _ in range(3):
    r = ctx.face_detect(casc, fr, 1920, 1.1, 3, (24, 24))
print(len(r), ctx.stage_times())
ctx.close()
