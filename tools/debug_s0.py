import os, sys, collections
import numpy as np
ROOT = "/root/repo"
sys.path[:0] = [os.path.join(ROOT, "nubomedia-vca_b200", "python"), os.path.join(ROOT, "oracle")]
import nubovca as nv, oracle as O
from nubovca import synth
xml = os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml")
nc, oc = nv.Cascade(xml), O.Cascade(xml)
fr = synth.frame(640, 480, 4, 1)
ctx = nv.Context(0, 1920, 1080, debug=True)
got = ctx.face_detect(nc, fr, 640, 1.25, 3, None)
exp, eq = O.face_process(fr, oc, 640, 1.25, 3, None)
print("got", got.tolist(), "exp", exp.tolist(), ctx.counters())
for i, lv in enumerate(O.eval_pyramid(eq, oc, 1.25, (32, 24), keep_integrals=False)):
    d = ctx.depth_map(i); o = lv["depth"]
    bad = d != o
    pairs = collections.Counter(zip(d[bad].tolist(), o[bad].tolist()))
    print("level", i, d.shape, "ystep", lv["ystep"], "mismatches", int(bad.sum()), pairs.most_common(6))
    if bad.any():
        ys, xs = np.nonzero(bad)
        print("   first", list(zip(ys[:8].tolist(), xs[:8].tolist())), "rows with mismatches", len(set(ys.tolist())), "of", d.shape[0],
              "x range", xs.min(), xs.max())
