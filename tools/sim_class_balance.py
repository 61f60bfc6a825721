"""Class balance of the bulk cascade kernels, simulated on the CPU oracle's stage-exit depth maps of one BASELINE config-3
frame (no GPU needed; the oracle takes about a minute): for every bulk stage, how many classifier steps a tile shape costs
(a stage of a tile costs as many steps as its fullest bank class has windows alive) against packing the tile's alive windows
32 to a step.  This script counts fullest-class sums only; the figures quoted in DESIGN.md section 4 (4.30 / 3.83 / 3.49 M for the
ystep-1 levels, the first equal to the DFMA count ncu reports for k_cascade_classes<1>) come from the same depth maps with a
stage rounded up to whole rounds of the kernel's eight warps.  Source of the per-stage table in profiles/r2_summary.md section 8.

  python tools/sim_class_balance.py [cache.pkl]"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "nubomedia-vca_b200", "python")]

cache = sys.argv[1] if len(sys.argv) > 1 else "/tmp/depth_cfg3.pkl"
if os.path.exists(cache):
    D = pickle.load(open(cache, "rb"))
else:
    import oracle as O
    from nubovca import synth
    oc = O.Cascade(os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml"))
    _, eq = O.face_process(synth.frame(1920, 1080, 6, 3), oc, 1920, 1.1, 3, (24, 24))
    D = [(lv["ystep"], lv["depth"]) for lv in O.eval_pyramid(eq, oc, 1.1, (24, 24))]
    pickle.dump(D, open(cache, "wb"))

NS = [3, 16, 21, 39, 33, 44, 50, 51, 56, 71]          # weak classifiers of stages 0..9 of haarcascade_frontalface_alt.xml
PASS = 1


def alive_at(d, s):                                    # depth codes: 1 = passed every stage, -s = failed stage s
    return (d == PASS) | ((d <= -s) & (d > -100))


for ys, (th, tw) in [(2, (32, 64)), (2, (64, 64)), (2, (64, 128)), (1, (32, 64)), (1, (64, 64)), (1, (64, 128))]:
    rows = []
    for s in range(1, 10):
        steps = ideal = alive = 0
        for lys, d in D:
            if lys != ys:
                continue
            a = alive_at(d, s)
            ny, nx = a.shape
            py, px = (ny + th - 1) // th * th, (nx + tw - 1) // tw * tw
            p = np.zeros((py, px), bool)
            p[:ny, :nx] = a
            t = p.reshape(py // th, th, px // tw, tw).transpose(0, 2, 1, 3).reshape(-1, th, tw)
            ly, lx = np.mgrid[0:th, 0:tw]
            cls = (lx + 20 * ly) % 32                  # bank class of window (lx, ly): the tile pitch is 4 (mod 8) words
            per_class = np.stack([t[:, cls == c].sum(1) for c in range(32)], 1)
            steps += int(per_class.max(1).sum())
            ideal += int(np.ceil(per_class.sum(1) / 32).sum())
            alive += int(a.sum())
        rows.append((s, NS[s], alive, steps, ideal))
    tot, tid = sum(r[3] * r[1] for r in rows), sum(r[4] * r[1] for r in rows)
    print(f"ystep {ys}, tiles of {tw}x{th} windows: {tot / 1e6:.2f} M classifier steps, {tid / 1e6:.2f} M with perfect packing per tile")
    for s, n, al, st, idl in rows:
        print(f"    stage {s}: {n:3d} classifiers, {al:8d} windows alive, fullest-class sum {st:6d}, packed {idl:6d}, ratio {st / max(idl, 1):.2f}")
