"""4:2:0 ingest: nv_face_detect_yuv / nv_yuv2bgr take a decoder's I420 / YV12 / NV12 / NV21 planes.
Contract: the result equals the reference block (kmsfacedetect.cpp:805-811) on cv::cvtColor(frame, COLOR_YUV2BGR_<fmt>).
CPU part: the oracle's conversion against cv2 live and against tests/golden/yuv_golden.json (generated from cv2 by
tests/golden/make_golden.py).  GPU part: the CUDA path through the C ABI against the oracle, bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

import nubovca as nv
import oracle as O
from nubovca import synth

try:
    import cv2
    cv2.setNumThreads(1)
except Exception:  # pragma: no cover
    cv2 = None

HERE = os.path.dirname(os.path.abspath(__file__))
FACE_XML = "haarcascade_frontalface_alt.xml"
FMTS = ["I420", "YV12", "NV12", "NV21"]
GOLD = json.load(open(os.path.join(HERE, "golden", "yuv_golden.json")))["cases"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rects_equal(a, b):
    a = np.asarray(a, np.int32).reshape(-1, 4); b = np.asarray(b, np.int32).reshape(-1, 4)
    return a.shape == b.shape and bool((a == b).all())


def ora_bgr(buf, w, h, fmt):
    return O.yuv420_to_bgr(*O.yuv420_planes(buf, w, h, fmt), fmt=fmt)


def strided(planes, pad, rng):
    """The same planes inside wider, noise-filled rows (a decoder's aligned strides)."""
    out = []
    for p in planes:
        if p is None:
            out.append(None); continue
        big = rng.integers(0, 256, (p.shape[0], p.shape[1] + pad), dtype=np.uint8)
        big[:, :p.shape[1]] = p
        out.append(big[:, :p.shape[1]])
    return tuple(out)


# ------------------------------------------------------------------------------------------
# CPU: oracle pinned against cv2
# ------------------------------------------------------------------------------------------
@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("size", [(2, 2), (64, 48), (66, 50), (1280, 720)])
def test_oracle_yuv2bgr_vs_cv2(fmt, size):
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    buf = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)        # full range: exercises every saturation
    code = getattr(cv2, "COLOR_YUV2BGR_" + fmt)
    assert (ora_bgr(buf, w, h, fmt) == cv2.cvtColor(buf, code)).all()
    planes = strided(O.yuv420_planes(buf, w, h, fmt), 6, rng)
    assert (O.yuv420_to_bgr(*planes, fmt=fmt) == cv2.cvtColor(buf, code)).all()


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_oracle_yuv_golden(idx, cascade_dir):
    c = GOLD[idx]
    buf = synth.to_yuv420(synth.frame(c["W"], c["H"], c["k"], c["seed"]), c["fmt"])
    assert sha(buf) == c["yuv_sha"], "synthetic generator drifted"
    bgr = ora_bgr(buf, c["W"], c["H"], c["fmt"])
    assert sha(bgr) == c["bgr_sha"]
    casc = O.Cascade(os.path.join(cascade_dir, FACE_XML))
    r, eq = O.face_process(bgr, casc, c["width_to_process"], c["scale_factor"], c["min_neighbors"], tuple(c["min_size"]))
    assert sha(eq) == c["eq_sha"] and rects_equal(r, c["grouped"])


def test_synth_planes_match_oracle_planes():
    buf = synth.to_yuv420(synth.frame(64, 48, 1, 5), "YV12")
    for a, b in zip(synth.yuv420_planes(buf, 64, 48, "YV12"), O.yuv420_planes(buf, 64, 48, "YV12")):
        assert (a == b).all()


# ------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    c = nv.Context(0, 1920, 1080, debug=True)
    yield c
    c.close()


@pytest.fixture(scope="module")
def face(cascade_dir):
    p = os.path.join(cascade_dir, FACE_XML)
    return nv.Cascade(p), O.Cascade(p)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("size", [(2, 2), (66, 50), (640, 480), (1920, 1080)])
def test_gpu_yuv2bgr(ctx, fmt, size):
    w, h = size
    rng = np.random.default_rng(w + h)
    buf = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)
    exp = ora_bgr(buf, w, h, fmt)
    assert (ctx.yuv2bgr(synth.yuv420_planes(buf, w, h, fmt), fmt) == exp).all()             # one block of memory
    planes = strided(synth.yuv420_planes(buf, w, h, fmt), 10, rng)                          # scattered, padded planes
    assert (ctx.yuv2bgr(planes, fmt) == exp).all()


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_gpu_face_detect_yuv_golden(ctx, face, idx):
    ncasc, ocasc = face
    c = GOLD[idx]
    buf = synth.to_yuv420(synth.frame(c["W"], c["H"], c["k"], c["seed"]), c["fmt"])
    got = ctx.face_detect_yuv(ncasc, synth.yuv420_planes(buf, c["W"], c["H"], c["fmt"]), c["fmt"], c["width_to_process"],
                              c["scale_factor"], c["min_neighbors"], tuple(c["min_size"]))
    assert rects_equal(got, c["grouped"])
    assert sha(ctx.gray()) == c["eq_sha"]


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("case", [(640, 480, 160), (640, 480, 320), (640, 480, 640), (1280, 720, 640), (1280, 720, 500),
                                  (322, 242, 100), (322, 242, 322), (644, 484, 322), (1920, 1080, 1920)])
def test_gpu_face_detect_yuv_vs_oracle(ctx, face, fmt, case):
    """every resize mode (copy, 2x box, linear), random full-range chroma on top of a face frame, padded strides"""
    ncasc, ocasc = face
    w, h, w2p = case
    rng = np.random.default_rng(w2p)
    buf = synth.to_yuv420(synth.frame(w, h, 3, 11 + w2p), fmt).copy()
    buf[h:] = np.clip(buf[h:].astype(np.int16) + rng.integers(-90, 91, buf[h:].shape), 0, 255).astype(np.uint8)
    exp, eq = O.face_process(ora_bgr(buf, w, h, fmt), ocasc, w2p, 1.2, 2, None)
    planes = synth.yuv420_planes(buf, w, h, fmt)
    # contiguous planes and 16-byte padded rows take the vectorised kernels when the output width is a multiple of 4,
    # rows padded by 14 bytes (stride not a multiple of 4) and the other widths the byte-per-lane kernel
    for pl in (planes, strided(planes, 14, rng), strided(planes, 16, rng)):
        got = ctx.face_detect_yuv(ncasc, pl, fmt, w2p, 1.2, 2, None)
        assert (ctx.gray() == eq).all()
        assert rects_equal(got, exp)


@pytest.mark.gpu
def test_gpu_yuv_graph_replay_and_format_switch(face):
    """a non-debug context replays the per-frame graph; switching format or planes must not replay a stale graph"""
    ncasc, ocasc = face
    c = nv.Context(0, 1280, 720)
    try:
        w, h = 1280, 720
        for rep in range(3):
            for fmt in ("I420", "NV12", "I420", "NV21"):
                for seed in (1, 2, 2, 2, 3):
                    fr = synth.frame(w, h, 3, seed)
                    buf = synth.to_yuv420(fr, fmt)
                    exp, _ = O.face_process(ora_bgr(buf, w, h, fmt), ocasc, 640, 1.25, 3, None)
                    c.face_submit_yuv(ncasc, synth.yuv420_planes(buf, w, h, fmt), fmt, 640, 1.25, 3, None)
                    assert rects_equal(c.face_collect(), exp)
            exp, _ = O.face_process(fr, ocasc, 640, 1.25, 3, None)              # and back to BGR through the same context
            assert rects_equal(c.face_detect(ncasc, fr, 640, 1.25, 3, None), exp)
    finally:
        c.close()


@pytest.mark.gpu
def test_gpu_yuv_argument_errors(ctx, face):
    ncasc, _ = face
    buf = np.zeros((72, 48), np.uint8)
    with pytest.raises(nv.NuboError):                      # odd height
        ctx.face_detect_yuv(ncasc, (buf[:47], buf[48:60, :24], buf[60:72, :24]), "I420")
    big = np.zeros((2162 * 3 // 2, 3842), np.uint8)
    with pytest.raises(nv.NuboError):                      # larger than the context
        ctx.face_detect_yuv(ncasc, synth.yuv420_planes(big, 3842, 2162, "NV12"), "NV12")


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["I420", "NV12"])
def test_gpu_face_element_on_yuv_frames(cascade_dir, fmt):
    """nv_element_transform_frame_yuv: the nubofacedetector mirror (gating, tracking, message) on 4:2:0 buffers equals the
    oracle-backed restatement fed with cvtColor's BGR; the other elements refuse 4:2:0 frames."""
    from element_ref import FaceRef
    w, h = 640, 480
    base = synth.frame(w, h, 4, 1)
    rng = np.random.default_rng(5)
    frames = [np.clip(base.astype(np.int16) + rng.integers(-3, 4, base.shape, dtype=np.int16), 0, 255).astype(np.uint8)
              for _ in range(5)] + [np.full((h, w, 3), 90, np.uint8)] * 3
    e = nv.Element("nubofacedetector", 0, cascade_dir)
    ref = FaceRef(O.Cascade(os.path.join(cascade_dir, FACE_XML)))
    e.set("process-x-every-4-frames", 2); ref.p["x4"] = 2
    seen = 0
    for i, f in enumerate(frames):
        buf = synth.to_yuv420(f, fmt)
        msg, pushed, sig = e.process_yuv(synth.yuv420_planes(buf, w, h, fmt), fmt, pts_ns=i * 33_000_000)
        assert pushed and sig is None
        assert msg == ref.process(ora_bgr(buf, w, h, fmt)), i
        seen += len(msg)
    assert seen > 0
    e.close()
    with pytest.raises(nv.NuboError):                                  # odd geometry is refused, not crashed on
        e2 = nv.Element("nubofacedetector", 0, cascade_dir)
        try:
            buf = synth.to_yuv420(base, fmt)
            y, c1 = synth.yuv420_planes(buf, w, h, fmt)[:2]
            e2.process_yuv((y[:h - 1], c1), "NV12")
        finally:
            e2.close()


@pytest.mark.gpu
def test_gpu_yuv_planes_already_on_the_device(face):
    """on_device = 1: the planes are device pointers (an NVDEC surface / GstCudaMemory), no host copy at all"""
    import ctypes as C
    torch = pytest.importorskip("torch")
    ncasc, ocasc = face
    w, h = 1280, 720
    c = nv.Context(0, w, h)
    try:
        for fmt in ("NV12", "I420"):
            buf = synth.to_yuv420(synth.frame(w, h, 3, 1000), fmt)
            exp, _ = O.face_process(ora_bgr(buf, w, h, fmt), ocasc, 640, 1.25, 3, None)
            d = torch.from_numpy(buf.copy()).cuda()
            f = nv.YuvFrame()
            f.format, f.width, f.height, f.on_device = nv._FMT[fmt], w, h, 1
            f.plane[0], f.stride[0] = d.data_ptr(), w
            f.plane[1] = d.data_ptr() + w * h
            if fmt == "NV12":
                f.stride[1] = w
            else:
                f.stride[1] = f.stride[2] = w // 2
                f.plane[2] = d.data_ptr() + w * h + w * h // 4
            p = nv.Context._face_params(640, 1.25, 3, None)
            for _ in range(3):                                   # third call replays the graph
                assert nv._lib.nv_face_submit_yuv(c.handle, ncasc.handle, C.byref(f), C.byref(p)) == 0
                assert rects_equal(c.face_collect(), exp)
    finally:
        c.close()


def _bgra(bgr):
    out = np.empty(bgr.shape[:2] + (4,), np.uint8)
    out[..., :3] = bgr; out[..., 3] = 255
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FMTS)
def test_gpu_tracker_on_yuv_frames(fmt, cascade_dir):
    """nv_tracker_process_yuv and the nubotracker mirror on 4:2:0 buffers: the motion objects of every frame equal the
    oracle's on cvtColor's BGR frame (alpha added), including a switch back to BGRA input on the same context."""
    w, h = 640, 360
    frames = synth.tracker_sequence(w, h, 7, seed=9, noise=50)
    ctx = nv.Context(0, w, h)
    st = O.TrackerState(w, h)
    e = nv.Element("nubotracker", 0, cascade_dir)
    est = O.TrackerState(w, h)
    nobj = 0
    try:
        for i, f in enumerate(frames):
            ts = 40.0 * (i + 1)
            buf = synth.to_yuv420(np.ascontiguousarray(f[..., :3]), fmt)
            bgra = _bgra(ora_bgr(buf, w, h, fmt))
            planes = synth.yuv420_planes(buf, w, h, fmt)
            if i == 4:                                                 # one BGRA frame in between: same state, same kernels
                got = ctx.tracker_process(bgra, ts)
            else:
                got = ctx.tracker_process_yuv(planes, fmt, ts)
            exp, _, _ = st.process(bgra, ts)
            assert got.shape == exp.shape and (got == exp).all(), i
            nobj += len(exp)
            msg, pushed, _ = e.process_yuv(planes, fmt, pts_ns=int(ts) * 1_000_000)
            eexp, _, _ = est.process(bgra, ts)
            assert [list(m[2:]) for m in msg] == eexp.tolist() and not pushed, i
        assert nobj > 0
    finally:
        ctx.close(); e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,factory,files", [
    ("eye", "nuboeyedetector", ("haarcascade_mcs_righteye.xml", "haarcascade_mcs_lefteye.xml")),
    ("mouth", "nubomouthdetector", ("haarcascade_mcs_mouth.xml",)),
    ("nose", "nubonosedetector", ("haarcascade_mcs_nose.xml",)),
    ("ear", "nuboeardetector", ("haarcascade_mcs_rightear.xml", "haarcascade_mcs_leftear.xml"))])
def test_gpu_nested_elements_on_yuv_frames(tmp_path, cascade_dir, kind, factory, files):
    """The nested elements (face stage + ROI cascades + temporal logic) fed with NV12 / I420 buffers: every message equals
    the oracle-backed restatement's on cvtColor's BGR frame.  Stand-in feature models as in tests/test_gpu_elements.py."""
    import shutil
    from cascade_xml_util import permissive_cascade
    from element_ref import EarRef, FeatureRef
    d = str(tmp_path)
    shutil.copy(os.path.join(cascade_dir, FACE_XML), os.path.join(d, FACE_XML))
    shutil.copy(os.path.join(cascade_dir, FACE_XML), os.path.join(d, "haarcascade_profileface.xml"))
    sizes = {"eye": (18, 12), "mouth": (25, 15), "nose": (18, 15), "ear": (12, 20)}
    for i, name in enumerate(files):
        permissive_cascade(os.path.join(d, name), np.random.default_rng(i), *sizes[kind])
    oc = lambda n: O.Cascade(os.path.join(d, n))                                                # noqa: E731
    w, h = 1280, 720
    base = synth.frame(w, h, 3, 2, smin=0.4, smax=0.6)
    rng = np.random.default_rng(3)
    frames = [np.clip(base.astype(np.int16) + rng.integers(-3, 4, base.shape, dtype=np.int16), 0, 255).astype(np.uint8) for _ in range(4)]
    total = 0
    for fmt in ("NV12", "I420"):
        e = nv.Element(factory, 0, d)
        ref = EarRef(oc("haarcascade_profileface.xml"), oc(files[0]), oc(files[1])) if kind == "ear" else \
            FeatureRef(kind, oc(FACE_XML), *[oc(f) for f in files])
        e.set("view-" + {"eye": "eyes", "mouth": "mouths", "nose": "noses", "ear": "ears"}[kind], 1)   # ignored on 4:2:0 buffers
        for i, f in enumerate(frames):
            buf = synth.to_yuv420(f, fmt)
            msg, _, _ = e.process_yuv(synth.yuv420_planes(buf, w, h, fmt), fmt, pts_ns=i * 33_000_000)
            assert msg == ref.process(ora_bgr(buf, w, h, fmt)), (kind, fmt, i)
            total += sum(1 for m in msg if m[1] != "face")
        e.close()
    assert total > 0
