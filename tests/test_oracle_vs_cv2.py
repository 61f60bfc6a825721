"""Pins the CPU oracle (oracle/nubo_oracle.c) against cv2 4.13 — the library whose calls ARE the
reference's hot path (kmsfacedetect.cpp:805-811 etc.) — live where cv2 imports,
and against the committed fixtures of tests/golden/make_golden.py everywhere.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle as O
from cascade_xml_util import (random_cascade, random_general_model, random_int_cascade, random_lbp_cascade, write_cascade,
                              write_lbp_cascade, write_old_format)
from nubovca import synth

try:
    import cv2
    cv2.setNumThreads(1)
except Exception:  # pragma: no cover
    cv2 = None

needs_cv2 = pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rects_equal(a, b):
    a = np.asarray(a, np.int32).reshape(-1, 4); b = np.asarray(b, np.int32).reshape(-1, 4)
    return a.shape == b.shape and bool((a == b).all())


# ------------------------------------------------------------------------------------------
# golden fixtures (no cv2 needed)
# ------------------------------------------------------------------------------------------
def _golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def test_cfg3_full_size_golden(cascade_dir):
    """BASELINE config 3 at full size (1920x1080, 40 levels, 3.66 M windows): the oracle against the committed cv2 output."""
    c = _golden("cfg3_golden.json")["cases"][0]
    fr = synth.frame(c["W"], c["H"], c["k"], c["seed"])
    assert sha(fr) == c["frame_sha"], "synthetic generator drifted"
    eq = O.equalize_hist(O.bgr2gray(fr))
    assert sha(eq) == c["eq_sha"]
    casc = O.Cascade(os.path.join(cascade_dir, c["cascade"]))
    assert rects_equal(O.detect_multiscale(eq, casc, c["scale_factor"], 0, tuple(c["min_size"])), c["raw"])
    assert rects_equal(O.detect_multiscale(eq, casc, c["scale_factor"], c["min_neighbors"], tuple(c["min_size"])), c["grouped"])


@pytest.mark.parametrize("idx", range(3))
def test_lbp_golden(idx):
    """BOOST/LBP cascades: the oracle against the committed cv2 output on the committed random models."""
    c = _golden("lbp_golden.json")["cases"][idx]
    eq = O.equalize_hist(O.bgr2gray(synth.frame(c["W"], c["H"], c["k"], c["seed"])))
    assert sha(eq) == c["eq_sha"]
    casc = O.Cascade(os.path.join(HERE, "golden", c["cascade"]))
    raw = O.detect_multiscale(eq, casc, c["scale_factor"], 0)
    assert len(raw) == c["n_raw"] and sha(raw.astype(np.int32)) == c["raw_sha"] and rects_equal(raw[:200], c["raw_head"])
    assert rects_equal(O.detect_multiscale(eq, casc, c["scale_factor"], c["min_neighbors"]), c["grouped"])


@pytest.mark.parametrize("idx", range(8))
def test_face_golden(idx, cascade_dir):
    c = _golden("face_golden.json")["cases"][idx]
    fr = synth.frame(c["W"], c["H"], c["k"], c["seed"])
    assert sha(fr) == c["frame_sha"], "synthetic generator drifted"
    isc = c["W"] // c["width_to_process"]
    rows, cols = int(np.rint(c["H"] / isc)), int(np.rint(c["W"] / isc))
    aux = O.resize_linear(fr, cols, rows)
    assert sha(aux) == c["resized_sha"]
    gray = O.bgr2gray(aux)
    assert sha(gray) == c["gray_sha"]
    eq = O.equalize_hist(gray)
    assert sha(eq) == c["eq_sha"]
    casc = O.Cascade(os.path.join(cascade_dir, c["cascade"]))
    ms = tuple(c["min_size"])
    assert rects_equal(O.detect_multiscale(eq, casc, c["scale_factor"], 0, ms), c["raw"])
    assert rects_equal(O.detect_multiscale(eq, casc, c["scale_factor"], c["min_neighbors"], ms), c["grouped"])
    if c["cascade"] == "haarcascade_frontalface_alt.xml":
        r, geq = O.face_process(fr, casc, c["width_to_process"], c["scale_factor"], c["min_neighbors"], ms)
        assert sha(geq) == c["eq_sha"] and rects_equal(r, c["grouped"])


@pytest.mark.parametrize("idx", range(2))
def test_tracker_golden(idx):
    c = _golden("tracker_golden.json")["cases"][idx]
    frames = synth.tracker_sequence(c["W"], c["H"], c["nframes"], c["seed"], noise=c["noise"])
    st = O.TrackerState(c["W"], c["H"])
    for i, (f, g) in enumerate(zip(frames, c["frames"])):
        # literal MHI pipeline with an injected millisecond timestamp; no area filter / merge
        rects, nraw, _ = st.process(f, ts=33.3 * (i + 1), threshold=c["threshold"], min_area=-1,
                                    max_area=1 << 40, distance=0)
        assert sha(st.prev) == g["gray_sha"]
        assert nraw == len(g["rects"])
        assert rects_equal(rects, g["rects"]), f"frame {i}"


# ------------------------------------------------------------------------------------------
# live cv2 pins
# ------------------------------------------------------------------------------------------
@needs_cv2
def test_bgr2gray_live():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    assert (O.bgr2gray(img) == cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)).all()
    img4 = rng.integers(0, 256, (45, 67, 4), dtype=np.uint8)
    assert (O.bgr2gray(img4) == cv2.cvtColor(img4, cv2.COLOR_BGRA2GRAY)).all()
    # the tracker passes BGRA with the 3-channel code (gstnubotracker.cpp:356): alpha ignored
    assert (O.bgr2gray(img4) == cv2.cvtColor(img4[..., :3], cv2.COLOR_BGR2GRAY)).all()


@needs_cv2
@pytest.mark.parametrize("src", [(640, 480), (1280, 720), (333, 217)])
def test_resize_linear_live(src):
    rng = np.random.default_rng(1)
    sw, sh = src
    s3 = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    for (dw, dh) in [(160, 120), (213, 160), (320, 180), (500, 281), (sw // 2, sh // 2), (sw, sh),
                     (sw + 100, sh + 57), (sw * 2, sh * 2), (100, 300)]:
        for s in (s3, s3[..., 1].copy()):
            assert (O.resize_linear(s, dw, dh) == cv2.resize(s, (dw, dh), interpolation=cv2.INTER_LINEAR)).all(), \
                (src, dw, dh, s.ndim)


@needs_cv2
def test_equalize_hist_live():
    rng = np.random.default_rng(2)
    for t in range(5):
        g = rng.integers(10 * t, 256 - 20 * t, (120, 160), dtype=np.uint8)
        assert (O.equalize_hist(g) == cv2.equalizeHist(g)).all()
    g = np.full((10, 10), 7, np.uint8)
    assert (O.equalize_hist(g) == cv2.equalizeHist(g)).all()


@needs_cv2
@pytest.mark.parametrize("src", [(640, 480), (1920, 1080), (160, 120)])
def test_pyramid_exact_live(src):
    rng = np.random.default_rng(3)
    sw, sh = src
    g = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    f = 1.0
    for _ in range(45):
        f *= 1.1
        dw, dh = O.level_size(sw, sh, np.float32(f))
        if dw < 8 or dh < 8:
            break
        assert (O.resize_linear_exact(g, dw, dh) == cv2.resize(g, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)).all()
    up = O.resize_linear_exact(g, sw + 37, sh + 11)
    assert (up == cv2.resize(g, (sw + 37, sh + 11), interpolation=cv2.INTER_LINEAR_EXACT)).all()


@needs_cv2
def test_integral_live():
    rng = np.random.default_rng(4)
    g = rng.integers(0, 256, (1080, 1920), dtype=np.uint8)
    s, q = O.integral(g)
    s2, q2 = cv2.integral2(g, sdepth=cv2.CV_32S, sqdepth=cv2.CV_64F)
    assert (s == s2).all()
    assert (q == (q2.astype(np.uint64) & 0xFFFFFFFF).astype(np.uint32)).all()   # consumed modulo 2^32


@needs_cv2
@pytest.mark.parametrize("W,H,sf,ms", [(173, 131, 1.1, (0, 0)), (320, 240, 1.25, (0, 0)), (333, 211, 1.1, (24, 24)),
                                       (640, 360, 1.25, (32, 18)), (257, 199, 1.07, (0, 0)), (401, 300, 1.31, (30, 30)),
                                       (259, 259, 1.13, (0, 0)), (105, 105, 1.2, (0, 0))])
def test_window_enumeration_live(tmp_path, W, H, sf, ms):
    """Always-pass cascade: cv2 reports EVERY visited window, which exposes the scale list, level sizes
    (float division), ystep, the stripe row limit and candidate rounding (float product)."""
    p = str(tmp_path / "pass.xml")
    write_cascade(p, 20, 20, [(-1.0, [(0, 1e30, 1.0, 1.0)])], [[(2, 2, 10, 10, -1.0), (4, 4, 5, 5, 2.0)]])
    g = np.random.default_rng(5).integers(0, 256, (H, W), dtype=np.uint8)
    a = cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=sf, minNeighbors=0, minSize=ms)
    assert rects_equal(a, O.detect_multiscale(g, O.Cascade(p), sf, 0, ms))


@needs_cv2
def test_candidate_rounding_is_float_product(tmp_path):
    # x=50 at the third 1.1 scale: float(x*sc) rounds differently from the double product
    p = str(tmp_path / "pass.xml")
    write_cascade(p, 20, 20, [(-1.0, [(0, 1e30, 1.0, 1.0)])], [[(2, 2, 10, 10, -1.0), (4, 4, 5, 5, 2.0)]])
    g = np.random.default_rng(6).integers(0, 256, (38, 90), dtype=np.uint8)
    a = cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.1, minNeighbors=0)
    assert rects_equal(a, O.detect_multiscale(g, O.Cascade(p), 1.1, 0))


@needs_cv2
def test_variance_reject_live(tmp_path):
    p = str(tmp_path / "pass.xml")
    write_cascade(p, 20, 20, [(-1.0, [(0, 1e30, 1.0, 1.0)])], [[(2, 2, 10, 10, -1.0), (4, 4, 5, 5, 2.0)]])
    rng = np.random.default_rng(7)
    g = np.clip(128 + 6 * rng.standard_normal((200, 300)), 0, 255).astype(np.uint8)   # sigma < 10: rejected
    g[50:120, 80:200] = rng.integers(0, 256, (70, 120))
    a = cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.2, minNeighbors=0)
    b = O.detect_multiscale(g, O.Cascade(p), 1.2, 0)
    assert 0 < len(b) < 40000 and rects_equal(a, b)


@needs_cv2
def test_stage_sum_accumulates_in_double(tmp_path):
    # leaves 1e8, 1, -1e8: a float accumulator gives 0 (< 0.5, reject), a double one gives 1 (pass)
    p = str(tmp_path / "acc.xml")
    write_cascade(p, 20, 20, [(0.5, [(0, 1e30, 1e8, 1e8), (0, 1e30, 1.0, 1.0), (0, 1e30, -1e8, -1e8)])],
                  [[(2, 2, 10, 10, -1.0), (4, 4, 5, 5, 2.0)]])
    g = np.random.default_rng(8).integers(0, 256, (100, 100), dtype=np.uint8)
    a = cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.2, minNeighbors=0)
    b = O.detect_multiscale(g, O.Cascade(p), 1.2, 0)
    assert len(a) > 0 and rects_equal(a, b)


@needs_cv2
def test_feature_value_bit_exact(tmp_path):
    """Brackets cv2's float feature value between thr=v (reject) and thr=nextafter(v) (accept)."""
    rng = np.random.default_rng(11)
    p = str(tmp_path / "f.xml")
    tested = 0
    for it in range(60):
        g = rng.integers(0, 256, (20, 20), dtype=np.uint8)
        rects = []
        for k in range(2 + it % 2):
            x = int(rng.integers(0, 15)); y = int(rng.integers(0, 15))
            w = int(rng.integers(1, 20 - x + 1)); h = int(rng.integers(1, 20 - y + 1))
            rects.append((x, y, w, h, float(np.float32(rng.uniform(-3, 3))) if it % 3 else [-1., 2., 3.][k]))
        write_cascade(p, 20, 20, [(0.5, [(0, 0.0, 1.0, 0.0)])], [rects])
        s, q = O.integral(g)
        v = O.feature_value(O.Cascade(p), s, q, 0, 0, 0)
        if v is None:
            continue
        tested += 1
        res = []
        for T in (v, np.nextafter(v, np.float32(np.inf))):
            write_cascade(p, 20, 20, [(0.5, [(0, float(T), 1.0, 0.0)])], [rects])
            res.append(len(cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.5, minNeighbors=0)))
        assert res == [0, 1], (it, v, rects)
    assert tested > 40


@needs_cv2
def test_stage0_skip_rule_live(tmp_path):
    """a stage-0 reject skips the next window; a later-stage reject does not."""
    g = np.zeros((40, 200), np.uint8)
    rng = np.random.default_rng(12)
    g[:] = rng.integers(0, 256, g.shape)
    for seed in range(6):
        p = str(tmp_path / f"r{seed}.xml")
        random_cascade(p, np.random.default_rng(100 + seed), nstages=3, max_trees=3)
        a = cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.3, minNeighbors=0)
        assert rects_equal(a, O.detect_multiscale(g, O.Cascade(p), 1.3, 0)), seed


@needs_cv2
@pytest.mark.parametrize("seed", range(8))
def test_random_cascades_live(tmp_path, seed):
    rng = np.random.default_rng(200 + seed)
    p = str(tmp_path / "rand.xml")
    random_cascade(p, rng, nstages=int(rng.integers(2, 7)), max_trees=8)
    W, H = int(rng.integers(60, 400)), int(rng.integers(60, 300))
    g = synth.frame(W, H, 2, seed)[..., 1] if seed % 2 else rng.integers(0, 256, (H, W), dtype=np.uint8)
    sf = float(rng.choice([1.1, 1.25, 1.4]))
    cc = cv2.CascadeClassifier(p); oc = O.Cascade(p)
    for mn in (0, 2):
        a = cc.detectMultiScale(g, scaleFactor=sf, minNeighbors=mn)
        assert rects_equal(a, O.detect_multiscale(g, oc, sf, mn)), (seed, mn)


@needs_cv2
@pytest.mark.parametrize("name", ["haarcascade_frontalface_alt.xml", "haarcascade_profileface.xml",
                                  "haarcascade_eye.xml", "haarcascade_frontalface_default.xml"])
def test_real_cascades_live(name, cascade_dir):
    path = os.path.join(cascade_dir, name)
    cc = cv2.CascadeClassifier(path); oc = O.Cascade(path)
    for (W, H, k, seed, sf, ms) in [(480, 360, 10, 5, 1.25, (0, 0)), (320, 240, 3, 7, 1.1, (24, 24))]:
        g = cv2.equalizeHist(cv2.cvtColor(synth.frame(W, H, k, seed), cv2.COLOR_BGR2GRAY))
        for mn in (0, 3):
            a = cc.detectMultiScale(g, scaleFactor=sf, minNeighbors=mn, minSize=ms)
            assert rects_equal(a, O.detect_multiscale(g, oc, sf, mn, ms)), (name, W, H, mn)


@pytest.mark.parametrize("idx", range(5))
def test_general_golden(idx, cascade_dir):
    """Tree / tilted models against the committed cv2 outputs (tests/golden/general_golden.json): no cv2 needed."""
    c = json.load(open(os.path.join(HERE, "golden", "general_golden.json")))["cases"][idx]
    eq = O.equalize_hist(O.bgr2gray(synth.frame(c["W"], c["H"], c["k"], c["seed"], smin=c["smin"], smax=c["smax"])))
    assert hashlib.sha256(eq.tobytes()).hexdigest() == c["eq_sha"]
    assert hashlib.sha256(O.integral_tilted(eq).tobytes()).hexdigest() == c["tilted_sha"]
    oc = O.Cascade(os.path.join(cascade_dir, c["cascade"]))
    ms = tuple(c["min_size"])
    assert rects_equal(O.detect_multiscale(eq, oc, c["scale_factor"], 0, ms), c["raw"])
    assert rects_equal(O.detect_multiscale(eq, oc, c["scale_factor"], c["min_neighbors"], ms), c["grouped"])
    assert len(c["raw"]) > 0


GENERAL = ["haarcascade_lefteye_2splits.xml", "haarcascade_righteye_2splits.xml", "haarcascade_smile.xml",
           "haarcascade_eye_tree_eyeglasses.xml", "haarcascade_frontalface_alt2.xml"]


@needs_cv2
def test_tilted_integral_live():
    """cv::integral's third output, the sums the tilted features read."""
    rng = np.random.default_rng(5)
    for (h, w) in [(1, 1), (2, 3), (5, 7), (37, 53), (90, 20), (20, 90), (120, 160)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert (O.integral_tilted(img) == cv2.integral3(img)[2]).all(), (h, w)
    img = np.full((64, 64), 255, np.uint8)
    assert (O.integral_tilted(img) == cv2.integral3(img)[2]).all()


@needs_cv2
@pytest.mark.parametrize("name", GENERAL)
def test_general_cascades_live(name, cascade_dir, tmp_path):
    """Tree weak classifiers and tilted features (predictOrdered): the eye / smile models the nested elements can
    load in place of the absent mcs_* files, in the new and in the OpenCV-2.x XML layout."""
    path = os.path.join(cascade_dir, name)
    old = str(tmp_path / "old.xml")
    write_old_format(old, O.parse_cascade_xml(path))
    cc, oc, cco, oco = cv2.CascadeClassifier(path), O.Cascade(path), cv2.CascadeClassifier(old), O.Cascade(old)
    assert oc.general and not cco.empty()
    n = 0
    for (W, H, k, seed, sf, ms) in [(480, 360, 10, 5, 1.25, (0, 0)), (320, 240, 3, 7, 1.1, (24, 24))]:
        g = cv2.equalizeHist(cv2.cvtColor(synth.frame(W, H, k, seed, smin=0.3, smax=0.6), cv2.COLOR_BGR2GRAY))
        for mn in (0, 2):
            a = cc.detectMultiScale(g, scaleFactor=sf, minNeighbors=mn, minSize=ms)
            assert rects_equal(a, O.detect_multiscale(g, oc, sf, mn, ms)), (name, W, H, mn)
            assert rects_equal(cco.detectMultiScale(g, scaleFactor=sf, minNeighbors=mn, minSize=ms),
                               O.detect_multiscale(g, oco, sf, mn, ms)), (name, "old layout", mn)
            assert rects_equal(a, O.detect_multiscale(g, oco, sf, mn, ms))
            n += len(a)
    assert n > 0, "the test images never fire this cascade"


@needs_cv2
@pytest.mark.parametrize("seed", range(6))
def test_random_trainer_shaped_and_general_cascades_live(tmp_path, seed):
    """(1) random stump cascades shaped like the trainer's output (-1 / 2 / 3 weights, half, third and checkerboard
    rects, up to 40 classifiers per stage) — what the exact-integer GPU kernels certify; (2) random cascades with tree
    weak classifiers and tilted features.  Both against cv2."""
    rng = np.random.default_rng(500 + seed)
    p = str(tmp_path / "int.xml")
    random_int_cascade(p, rng, w=[20, 24, 18, 32, 20, 25][seed], h=[20, 24, 15, 20, 30, 15][seed])
    g = synth.frame(300, 220, 3, seed)[..., 1]
    for mn in (0, 2):
        assert rects_equal(cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.2, minNeighbors=mn),
                           O.detect_multiscale(g, O.Cascade(p), 1.2, mn)), (seed, mn)
    q = str(tmp_path / "gen.xml")
    write_old_format(q, random_general_model(rng))
    cc = cv2.CascadeClassifier(q)
    assert not cc.empty() and O.Cascade(q).general
    for mn in (0, 2):
        assert rects_equal(cc.detectMultiScale(g, scaleFactor=1.2, minNeighbors=mn), O.detect_multiscale(g, O.Cascade(q), 1.2, mn)), (seed, mn)


# ---- LBP cascades.  Neither the reference nor this image ships an LBP model, so the oracle's LBP
# path is pinned on purpose-built and random models that cv2 loads.
@needs_cv2
def test_lbp_stage_sum_accumulates_in_double(tmp_path):
    # leaves 2^24, 1, 1, 1, 1 against a threshold of 2^24 + 2: a float accumulator stays at 2^24 (reject)
    A = 16777216.0
    z = [0] * 8
    trees = [([(0, -1, 0, z)], [A, A])] + [([(0, -1, 0, z)], [1.0, 1.0]) for _ in range(4)]
    p = str(tmp_path / "acc.xml")
    write_lbp_cascade(p, 24, 24, [(A + 2.0, trees)], [(0, 0, 3, 3)])
    g = np.random.default_rng(8).integers(0, 256, (80, 90), dtype=np.uint8)
    a = cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.2, minNeighbors=0)
    b = O.detect_multiscale(g, O.Cascade(p), 1.2, 0)
    assert len(a) > 0 and rects_equal(a, b)


@needs_cv2
def test_lbp_code_bit_exact(tmp_path):
    """The 8-bit code of one feature, read out of cv2 through single-code subsets: with subset = {code} the window
    passes, with subset = all codes but that one it does not — for random cells on random 24x24 images (one window)."""
    rng = np.random.default_rng(21)
    p = str(tmp_path / "code.xml")
    seen = set()
    for it in range(80):
        g = rng.integers(0, 256, (24, 24), dtype=np.uint8) if it % 4 else np.full((24, 24), it, np.uint8)   # flat image: code 255
        if it % 4 == 1:
            g = (g // 64 * 64).astype(np.uint8)                      # coarse values: ties between cells are common
        cw, ch = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        cell = (int(rng.integers(0, 24 - 3 * cw + 1)), int(rng.integers(0, 24 - 3 * ch + 1)), cw, ch)
        write_lbp_cascade(p, 24, 24, [(0.5, [([(0, -1, 0, [0] * 8)], [1.0, 0.0])])], [cell])
        s, _ = O.integral(g)
        code = O.lbp_code(O.Cascade(p), s, 0, 0, 0)
        seen.add(code)
        only = [0] * 8; only[code >> 5] = 1 << (code & 31)
        if only[code >> 5] >= 2**31:
            only[code >> 5] -= 2**32
        rest = [(-1 if i != code >> 5 else (~(1 << (code & 31))) & 0xffffffff) for i in range(8)]
        rest = [v - 2**32 if v >= 2**31 else v for v in rest]
        res = []
        for sub in (only, rest):
            write_lbp_cascade(p, 24, 24, [(0.5, [([(0, -1, 0, sub)], [1.0, 0.0])])], [cell])
            res.append(len(cv2.CascadeClassifier(p).detectMultiScale(g, scaleFactor=1.5, minNeighbors=0)))
        assert res == [1, 0], (it, code, cell)
    assert len(seen) > 20 and 255 in seen


@needs_cv2
@pytest.mark.parametrize("seed", range(8))
def test_random_lbp_cascades_live(tmp_path, seed):
    """Random LBP cascades — stumps (predictCategoricalStump) for even seeds, trees of up to three nodes
    (predictCategorical) for odd ones — on noise and on synthetic frames, raw and grouped, against cv2."""
    rng = np.random.default_rng(700 + seed)
    p = str(tmp_path / "lbp.xml")
    w, h = [(24, 24), (20, 20), (32, 18), (18, 30)][seed % 4]
    random_lbp_cascade(p, rng, w=w, h=h, nstages=int(rng.integers(2, 8)), max_trees=8, max_nodes=1 if seed % 2 == 0 else 3)
    W, H = int(rng.integers(60, 400)), int(rng.integers(60, 300))
    g = synth.frame(W, H, 2, seed)[..., 1] if seed % 3 else rng.integers(0, 256, (H, W), dtype=np.uint8)
    sf = float(rng.choice([1.1, 1.25, 1.4]))
    cc = cv2.CascadeClassifier(p); oc = O.Cascade(p)
    assert not cc.empty() and oc.lbp
    n = 0
    for mn in (0, 2):
        a = cc.detectMultiScale(g, scaleFactor=sf, minNeighbors=mn)
        assert rects_equal(a, O.detect_multiscale(g, oc, sf, mn)), (seed, mn)
        n += len(a)
    assert n > 0, "the random model never fires on this image"


@needs_cv2
def test_old_format_cascade_live(tmp_path, cascade_dir):
    """cv2 4.13 converts OpenCV-2.x "opencv-haar-classifier" files on load and evaluates them exactly like the
    new layout, whatever the flags: the old-format parser must therefore give the same detections."""
    src = os.path.join(cascade_dir, "haarcascade_frontalface_alt.xml")
    p = str(tmp_path / "old.xml")
    write_old_format(p, O.parse_cascade_xml(src))
    g = cv2.equalizeHist(cv2.cvtColor(synth.frame(480, 360, 10, 5), cv2.COLOR_BGR2GRAY))
    cc = cv2.CascadeClassifier(p)
    assert not cc.empty()
    exp = O.detect_multiscale(g, O.Cascade(p), 1.25, 0)
    assert len(exp) > 10 and rects_equal(exp, O.detect_multiscale(g, O.Cascade(src), 1.25, 0))
    for flags in (0, cv2.CASCADE_SCALE_IMAGE, cv2.CASCADE_FIND_BIGGEST_OBJECT):
        assert rects_equal(cc.detectMultiScale(g, scaleFactor=1.25, minNeighbors=0, flags=flags), exp)


@needs_cv2
def test_shipped_old_format_file_live():
    """haarcascade_license_plate_rus_16stages.xml is a genuine old-layout file with a 64x16 window."""
    p = os.path.join(cv2.data.haarcascades, "haarcascade_license_plate_rus_16stages.xml")
    cc = cv2.CascadeClassifier(p); oc = O.Cascade(p)
    rng = np.random.default_rng(21)
    g = cv2.equalizeHist(cv2.GaussianBlur(rng.integers(0, 256, (240, 400), dtype=np.uint8), (0, 0), 1.2))
    n = 0
    for mn in (0, 2):
        a = cc.detectMultiScale(g, scaleFactor=1.1, minNeighbors=mn)
        assert rects_equal(a, O.detect_multiscale(g, oc, 1.1, mn))
        n += len(a)
    assert n > 0


@needs_cv2
def test_multithreaded_cv2_gives_same_set(cascade_dir):
    path = os.path.join(cascade_dir, "haarcascade_frontalface_alt.xml")
    g = cv2.equalizeHist(cv2.cvtColor(synth.frame(480, 360, 10, 5), cv2.COLOR_BGR2GRAY))
    cv2.setNumThreads(4)
    try:
        a = cv2.CascadeClassifier(path).detectMultiScale(g, scaleFactor=1.25, minNeighbors=0)
    finally:
        cv2.setNumThreads(1)
    b = O.detect_multiscale(g, O.Cascade(path), 1.25, 0)
    assert sorted(map(tuple, np.asarray(a).reshape(-1, 4).tolist())) == sorted(map(tuple, b.tolist()))


@needs_cv2
def test_group_rectangles_live():
    rng = np.random.default_rng(13)
    for it in range(200):
        n = int(rng.integers(0, 60))
        base = rng.integers(0, 200, (max(1, n // 4), 2))
        r = []
        for _ in range(n):
            b = base[rng.integers(0, len(base))]
            s = int(rng.integers(20, 90))
            r.append([int(b[0] + rng.integers(-6, 7)), int(b[1] + rng.integers(-6, 7)),
                      s + int(rng.integers(-3, 4)), s + int(rng.integers(-3, 4))])
        thr = int(rng.integers(0, 4))
        a, w = cv2.groupRectangles([list(x) for x in r], thr, 0.2) if n else ([], [])
        b, wb = O.group_rectangles(r, thr)
        assert rects_equal(a, b), it
        if thr > 0 and n:
            assert list(np.asarray(w).reshape(-1)) == list(wb)


@needs_cv2
def test_segment_motion_matches_floodfill_emulation():
    """cv2.motempl is absent: emulate segmentMotion's outer loop with the real cv2.floodFill
    (4-connected, floating range lo=up=32, mask only) and compare rects + order."""
    frames = synth.tracker_sequence(320, 180, 5, seed=21, noise=60)
    prev = None
    for i, f in enumerate(frames):
        gray = cv2.cvtColor(f, cv2.COLOR_BGRA2GRAY)
        if prev is not None:
            ts = 40.0 * i
            mask = cv2.threshold(cv2.absdiff(gray, prev), 20, 255, cv2.THRESH_BINARY)[1]
            mhi = np.where(mask > 0, np.float32(ts), np.float32(0))
            work = np.where(mhi == 0, np.float32(3.4028234e37), mhi).astype(np.float32)
            ffmask = np.zeros((182, 322), np.uint8)
            exp = []
            for y in range(180):
                for x in np.flatnonzero((work[y] == np.float32(ts)) & (ffmask[y + 1, 1:-1] == 0)):
                    if ffmask[y + 1, x + 1]:
                        continue
                    _, _, _, rect = cv2.floodFill(work, ffmask, (int(x), y), 0, 32, 32,
                                                  cv2.FLOODFILL_MASK_ONLY | (2 << 8) | 4)
                    ffmask[ffmask == 2] = 1
                    exp.append(list(rect))
            got, _ = O.segment_motion(mhi, ts)
            assert rects_equal(got, exp), i
        prev = gray


def test_mhi_update_literal():
    """updateMotionHistory with ms timestamps and duration 0.2 leaves mhi == ts * (mask != 0)
    (gstnubotracker.cpp:28,349,368): stale entries are always older than ts - 0.2."""
    import ctypes as C
    rng = np.random.default_rng(14)
    mhi = np.zeros(1000, np.float32)
    for i in range(1, 5):
        silh = (rng.random(1000) < 0.3).astype(np.uint8) * 255
        O.lib().ora_update_mhi(silh.ctypes.data_as(C.c_void_p), mhi.ctypes.data_as(C.c_void_p), 1000, 33.3 * i, 0.2)
        assert (mhi == np.where(silh > 0, np.float32(33.3 * i), np.float32(0))).all()


def test_join_objects_cases():
    # area filter is exclusive on both sides; merge is back-to-front (gstnubotracker.cpp:171-200)
    r = [[0, 0, 10, 5], [100, 100, 10, 10], [104, 104, 10, 10], [300, 300, 200, 200], [0, 0, 7, 7]]
    out = O.join_objects(r, 50, 30000, 35)
    assert out.tolist() == [[100, 100, 14, 14]]
    # containment keeps the outer rectangle
    out = O.join_objects([[10, 10, 50, 50], [20, 20, 10, 10]], 50, 30000, 35)
    assert out.tolist() == [[10, 10, 50, 50]]
    assert O.join_objects([], 50, 30000, 35).shape == (0, 4)
