"""Generates tests/golden/*.json from cv2 4.13 (the library the reference calls for every
arithmetic step of the hot path).  Run here, in the build container:

    python tests/golden/make_golden.py

The fixtures hold, for seeded synthetic frames (nubovca.synth): sha256 of cv2's gray /
resized / equalised images and cv2's detectMultiScale rectangles (single-thread order) for
the element parameter sets of BASELINE.json configs, plus tracker component rectangles from
cv2.connectedComponentsWithStats.  Tests compare the oracle (and, on the GPU box, the CUDA
path) with these files without needing cv2.
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
from nubovca import synth  # noqa: E402

CASC = os.path.join(ROOT, "nubomedia-vca_b200", "cascades")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def face_case(W, H, k, seed, w2p, sf, mn, min_size, cascade="haarcascade_frontalface_alt.xml"):
    """kmsfacedetect.cpp:770-811 through cv2."""
    fr = synth.frame(W, H, k, seed)
    iscale = W // w2p
    rows, cols = int(np.rint(H / iscale)), int(np.rint(W / iscale))
    aux = cv2.resize(fr, (cols, rows), interpolation=cv2.INTER_LINEAR)
    gray = cv2.cvtColor(aux, cv2.COLOR_BGR2GRAY)
    eq = cv2.equalizeHist(gray)
    ms = (cols // 20, rows // 20) if min_size is None else tuple(min_size)
    cc = cv2.CascadeClassifier(os.path.join(CASC, cascade))
    raw = np.asarray(cc.detectMultiScale(eq, scaleFactor=sf, minNeighbors=0, minSize=ms)).reshape(-1, 4)
    grp = np.asarray(cc.detectMultiScale(eq, scaleFactor=sf, minNeighbors=mn, minSize=ms)).reshape(-1, 4)
    return dict(W=W, H=H, k=k, seed=seed, width_to_process=w2p, scale_factor=sf, min_neighbors=mn,
                min_size=list(ms), cascade=cascade, frame_sha=sha(fr), resized_sha=sha(aux), gray_sha=sha(gray),
                eq_sha=sha(eq), raw=raw.tolist(), grouped=grp.tolist())


def tracker_case(W, H, nframes, seed, thr, noise):
    frames = synth.tracker_sequence(W, H, nframes, seed, noise=noise)
    prev = None
    out = []
    for f in frames:
        gray = cv2.cvtColor(f, cv2.COLOR_BGRA2GRAY)
        rects = []
        if prev is not None:
            mask = cv2.threshold(cv2.absdiff(gray, prev), thr, 255, cv2.THRESH_BINARY)[1]
            n, lab, stats, _ = cv2.connectedComponentsWithStats(mask, connectivity=4)
            first = {}
            flat = lab.reshape(-1)
            idx = np.flatnonzero(flat)
            for i in idx:   # raster order of first pixel
                first.setdefault(int(flat[i]), int(i))
            for l in sorted(first, key=first.get):
                rects.append([int(stats[l, 0]), int(stats[l, 1]), int(stats[l, 2]), int(stats[l, 3])])
        out.append(dict(gray_sha=sha(gray), rects=rects))
        prev = gray
    return dict(W=W, H=H, nframes=nframes, seed=seed, threshold=thr, noise=noise, frames=out)


def general_case(cascade, W, H, k, seed, smin, smax, sf, mn, min_size):
    """cv::CascadeClassifier::detectMultiScale with a tree / tilted model (predictOrdered) on an equalised gray frame,
    plus the sha256 of cv2.integral3's tilted sums of that frame."""
    eq = cv2.equalizeHist(cv2.cvtColor(synth.frame(W, H, k, seed, smin=smin, smax=smax), cv2.COLOR_BGR2GRAY))
    cc = cv2.CascadeClassifier(os.path.join(CASC, cascade))
    raw = np.asarray(cc.detectMultiScale(eq, scaleFactor=sf, minNeighbors=0, minSize=tuple(min_size))).reshape(-1, 4)
    grp = np.asarray(cc.detectMultiScale(eq, scaleFactor=sf, minNeighbors=mn, minSize=tuple(min_size))).reshape(-1, 4)
    return dict(cascade=cascade, W=W, H=H, k=k, seed=seed, smin=smin, smax=smax, scale_factor=sf, min_neighbors=mn,
                min_size=list(min_size), eq_sha=sha(eq), tilted_sha=sha(cv2.integral3(eq)[2].astype(np.int32)),
                raw=raw.tolist(), grouped=grp.tolist())


YUV_CODES = {"I420": cv2.COLOR_YUV2BGR_I420, "YV12": cv2.COLOR_YUV2BGR_YV12, "NV12": cv2.COLOR_YUV2BGR_NV12,
             "NV21": cv2.COLOR_YUV2BGR_NV21}


def yuv_case(fmt, W, H, k, seed, w2p, sf, mn, min_size):
    """4:2:0 ingest: cvtColor(COLOR_YUV2BGR_<fmt>) in front of the face block (kmsfacedetect.cpp:805-811)."""
    buf = synth.to_yuv420(synth.frame(W, H, k, seed), fmt)
    bgr = cv2.cvtColor(buf, YUV_CODES[fmt])
    iscale = W // w2p
    rows, cols = int(np.rint(H / iscale)), int(np.rint(W / iscale))
    eq = cv2.equalizeHist(cv2.cvtColor(cv2.resize(bgr, (cols, rows), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY))
    ms = (cols // 20, rows // 20) if min_size is None else tuple(min_size)
    cc = cv2.CascadeClassifier(os.path.join(CASC, "haarcascade_frontalface_alt.xml"))
    grp = np.asarray(cc.detectMultiScale(eq, scaleFactor=sf, minNeighbors=mn, minSize=ms)).reshape(-1, 4)
    return dict(fmt=fmt, W=W, H=H, k=k, seed=seed, width_to_process=w2p, scale_factor=sf, min_neighbors=mn,
                min_size=list(ms), yuv_sha=sha(buf), bgr_sha=sha(bgr), eq_sha=sha(eq), grouped=grp.tolist())


def lbp_cases():
    """BOOST/LBP cascades (categorical stumps and trees).  No LBP model ships with this image or with the reference, so
    three random models are written next to this file (tests/cascade_xml_util.random_lbp_cascade, seeded) and cv2's
    detectMultiScale output on them is the fixture."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cascade_xml_util import random_lbp_cascade
    cv2.setNumThreads(1)                 # canonical candidate order: scale -> y -> x
    out = []
    for i, (seed, w, h, nst, max_nodes, W, H, sf) in enumerate([(41, 24, 24, 7, 1, 320, 240, 1.1), (42, 20, 20, 6, 3, 400, 300, 1.2),
                                                                (43, 32, 18, 8, 2, 640, 360, 1.1)]):
        name = f"lbp_random_{i}.xml"
        random_lbp_cascade(os.path.join(HERE, name), np.random.default_rng(seed), w=w, h=h, nstages=nst, max_trees=10, max_nodes=max_nodes,
                           pass_bias=0.02)
        eq = cv2.equalizeHist(cv2.cvtColor(synth.frame(W, H, 3, seed), cv2.COLOR_BGR2GRAY))
        cc = cv2.CascadeClassifier(os.path.join(HERE, name))
        assert not cc.empty()
        raw = np.asarray(cc.detectMultiScale(eq, scaleFactor=sf, minNeighbors=0)).reshape(-1, 4)
        grp = np.asarray(cc.detectMultiScale(eq, scaleFactor=sf, minNeighbors=2)).reshape(-1, 4)
        print(name, len(raw), "raw", len(grp), "grouped")
        out.append(dict(cascade=name, W=W, H=H, k=3, seed=seed, scale_factor=sf, min_neighbors=2, eq_sha=sha(eq),
                        raw_sha=sha(raw.astype(np.int32)), n_raw=len(raw), raw_head=raw[:200].tolist(), grouped=grp.tolist()))
    with open(os.path.join(HERE, "lbp_golden.json"), "w") as f:
        json.dump(dict(cv2=cv2.__version__, cases=out), f)
    print("wrote", len(out), "LBP cascade cases")


def cfg3_full():
    """BASELINE config 3 at FULL size (1920x1080, processing width 1920, sf 1.1, min 24x24, the bench.py frames of rank 0):
    its own file, so that regenerating it does not touch the other fixtures."""
    cases = [face_case(1920, 1080, 6, 3 + i, 1920, 1.1, 3, (24, 24)) for i in range(2)]
    with open(os.path.join(HERE, "cfg3_golden.json"), "w") as f:
        json.dump(dict(cv2=cv2.__version__, cases=cases), f)
    print("wrote", len(cases), "full-size config-3 cases:", [(len(c["raw"]), len(c["grouped"])) for c in cases])


def main():
    cv2.setNumThreads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "--cfg3":
        return cfg3_full()
    yuv = [yuv_case("I420", 640, 480, 4, 1, 160, 1.25, 3, None),            # linear resize (4x)
           yuv_case("NV12", 1280, 720, 3, 1000, 640, 1.25, 3, None),         # cfg5 stream 0: the 2x box path
           yuv_case("NV21", 640, 360, 6, 3, 640, 1.1, 3, (24, 24)),          # no resize
           yuv_case("YV12", 642, 362, 4, 7, 214, 1.2, 2, (0, 0))]            # odd chroma width, 3x
    with open(os.path.join(HERE, "yuv_golden.json"), "w") as f:
        json.dump(dict(cv2=cv2.__version__, cases=yuv), f)
    print("wrote", len(yuv), "4:2:0 ingest cases")
    gen = [general_case(c, 400, 300, 2, 3, 0.5, 0.9, sf, 2, ms) for c, sf, ms in [
        ("haarcascade_lefteye_2splits.xml", 1.1, (20, 20)), ("haarcascade_righteye_2splits.xml", 1.1, (0, 0)),
        ("haarcascade_smile.xml", 1.1, (1, 1)), ("haarcascade_eye_tree_eyeglasses.xml", 1.25, (0, 0)),
        ("haarcascade_frontalface_alt2.xml", 1.2, (0, 0))]]
    with open(os.path.join(HERE, "general_golden.json"), "w") as f:
        json.dump(dict(cv2=cv2.__version__, cases=gen), f)
    print("wrote", len(gen), "tree / tilted cascade cases")
    faces = [
        face_case(640, 480, 4, 1, 160, 1.25, 3, None),                       # cfg1: element defaults
        face_case(640, 480, 4, 1, 640, 1.25, 3, None),
        face_case(1280, 720, 3, 1000, 640, 1.25, 3, None),                   # cfg5 stream 0
        face_case(1280, 720, 5, 2, 160, 1.25, 3, (30, 30)),                  # cfg2 face stage (eye element)
        face_case(1280, 720, 5, 2, 160, 1.25, 2, (3, 3)),                    # cfg2 face stage (mouth/nose)
        face_case(640, 360, 6, 3, 640, 1.1, 3, (24, 24)),                    # cfg3 parameters, reduced size
        face_case(640, 480, 4, 9, 320, 1.1, 2, (0, 0), "haarcascade_profileface.xml"),
        face_case(640, 480, 4, 9, 320, 1.1, 2, (20, 20), "haarcascade_eye.xml"),
    ]
    with open(os.path.join(HERE, "face_golden.json"), "w") as f:
        json.dump(dict(cv2=cv2.__version__, cases=faces), f)
    trk = [tracker_case(320, 180, 6, 4, 20, 40), tracker_case(640, 360, 4, 5, 20, 0)]
    with open(os.path.join(HERE, "tracker_golden.json"), "w") as f:
        json.dump(dict(cv2=cv2.__version__, cases=trk), f)
    print("wrote", len(faces), "face cases,", len(trk), "tracker cases")


if __name__ == "__main__":
    if sys.argv[1:] == ["lbp"]:          # only the LBP fixtures (added in round 2; the others are unchanged)
        lbp_cases()
    else:
        main()
        lbp_cases()
