"""Randomised parity run of the element mirrors (not part of the test suite): nubofacedetector, nuboeyedetector,
nubomouthdetector, nubonosedetector and nuboeardetector with random frame sizes, face layouts, property values
(width-to-process, multi-scale-factor, process-x-every-4-frames) and random stand-in feature models, frame sequences with
sensor noise and empty frames in between (temporal logic), against the oracle-backed restatement tests/element_ref.py,
message by message.  Usage: python tools/fuzz_elements.py [seconds] [seed]"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "nubomedia-vca_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import nubovca as nv  # noqa: E402
import oracle as O  # noqa: E402
from cascade_xml_util import permissive_cascade  # noqa: E402
from element_ref import EarRef, FaceRef, FeatureRef  # noqa: E402
from nubovca import synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
SRC = os.path.join(ROOT, "nubomedia-vca_b200", "cascades")
STANDINS = {"haarcascade_mcs_righteye.xml": (18, 12), "haarcascade_mcs_lefteye.xml": (18, 12), "haarcascade_mcs_mouth.xml": (25, 15),
            "haarcascade_mcs_nose.xml": (18, 15), "haarcascade_mcs_rightear.xml": (12, 20), "haarcascade_mcs_leftear.xml": (12, 20)}
KINDS = [("face", "nubofacedetector", ()), ("eye", "nuboeyedetector", ("haarcascade_mcs_righteye.xml", "haarcascade_mcs_lefteye.xml")),
         ("mouth", "nubomouthdetector", ("haarcascade_mcs_mouth.xml",)), ("nose", "nubonosedetector", ("haarcascade_mcs_nose.xml",)),
         ("ear", "nuboeardetector", ("haarcascade_mcs_rightear.xml", "haarcascade_mcs_leftear.xml"))]
t0 = time.time()
nseq = nframes = nfeat = bad = 0
while time.time() - t0 < budget:
    d = tempfile.mkdtemp(prefix="nubovca_fz_")
    try:
        shutil.copy(os.path.join(SRC, "haarcascade_frontalface_alt.xml"), d)
        shutil.copy(os.path.join(SRC, "haarcascade_frontalface_alt.xml"), os.path.join(d, "haarcascade_profileface.xml"))
        for name, (w, h) in STANDINS.items():
            permissive_cascade(os.path.join(d, name), np.random.default_rng(int(rng.integers(1 << 30))), w, h,
                               bias=float(rng.choice([0.2, 0.35, 0.5])))
        oc = lambda n: O.Cascade(os.path.join(d, n))                                            # noqa: E731
        kind, factory, files = KINDS[int(rng.integers(len(KINDS)))]
        W, H = [(640, 480), (1280, 720), (960, 540), (800, 600), (1920, 1080)][int(rng.integers(5))]
        base = synth.frame(W, H, int(rng.integers(1, 4)), int(rng.integers(1 << 30)), smin=0.3, smax=0.7)
        frames = []
        for i in range(int(rng.integers(3, 8))):
            if rng.random() < 0.2:
                frames.append(np.full((H, W, 3), 90, np.uint8))                                  # nothing to detect
            else:
                frames.append(np.clip(base.astype(np.int16) + rng.integers(-3, 4, base.shape, dtype=np.int16), 0, 255).astype(np.uint8))
        e = nv.Element(factory, 0, d)
        if kind == "face":
            ref = FaceRef(oc("haarcascade_frontalface_alt.xml"))
        elif kind == "ear":
            ref = EarRef(oc("haarcascade_profileface.xml"), oc(files[0]), oc(files[1]))
        else:
            ref = FeatureRef(kind, oc("haarcascade_frontalface_alt.xml"), *[oc(f) for f in files])
        x4 = int(rng.choice([1, 2, 3, 4])); sf = int(rng.choice([10, 25, 40]))
        w2p = int(rng.choice([160, 320] if kind == "face" else [320, 640, 480]))
        e.set("process-x-every-4-frames", x4); ref.p["x4"] = x4
        e.set("multi-scale-factor", sf); ref.p["sf"] = sf
        e.set("width-to-process", w2p); ref.p["w2p"] = w2p
        for i, f in enumerate(frames):
            msg, _, _ = e.process(f, pts_ns=i * 33_000_000)
            exp = ref.process(f)
            nframes += 1
            nfeat += sum(1 for m in msg if m[1] != "face")
            if msg != exp:
                bad += 1
                print("MISMATCH", kind, W, H, x4, sf, w2p, i, msg[:3], exp[:3], flush=True)
                break
        e.close()
        nseq += 1
    finally:
        shutil.rmtree(d, ignore_errors=True)
print(f"fuzz_elements: {nseq} sequences, {nframes} frames, {nfeat} feature rectangles, {bad} mismatches, seed {seed}, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
