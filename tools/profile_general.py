"""Where the time of a tree / tilted model on a large image goes: per-stage CUDA-event times and the window / candidate
counters of one call (haarcascade_smile.xml and frontalface_alt2 on a 1920x1080 frame)."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nubovca as nv
import oracle as O
from nubovca import synth
c = nv.Context(0, 1920, 1080)
c.set_profile(True)
cd = os.path.join(ROOT, "nubomedia-vca_b200", "cascades")
g = O.equalize_hist(O.bgr2gray(synth.frame(1920, 1080, 4, 9, smin=0.2, smax=0.5)))
for name in ["haarcascade_smile.xml", "haarcascade_frontalface_alt2.xml", "haarcascade_lefteye_2splits.xml"]:
    nc = nv.Cascade(os.path.join(cd, name))
    for mn in (3, 0):
        for _ in range(2):
            r = c.detect_multiscale(nc, g, 1.1, mn)
        t = time.perf_counter(); r = c.detect_multiscale(nc, g, 1.1, mn); dt = time.perf_counter() - t
        print(name, "minNeighbors", mn, "rects", len(r), "wall ms %.3f" % (dt * 1e3), {k: round(v, 3) for k, v in c.stage_times().items()}, c.counters())
c.close()
