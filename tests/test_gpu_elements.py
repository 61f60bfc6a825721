"""GPU tests of the element mirrors (BASELINE config 2: nested facial features inside face ROIs, plus the
face element over a frame sequence): the CUDA-backed elements against the oracle-backed restatement
(tests/element_ref.py) on the same synthetic frames, message by message.

The six haarcascade_mcs_*.xml files the feature elements load are absent from this image,
so the cascade directory of the test holds stand-ins under the reference's file names: random stump cascades
with the real models' window sizes that fire often (content-agnostic, never vacuous)."""
import os
import shutil

import numpy as np
import pytest

import nubovca as nv
import oracle as O
from cascade_xml_util import permissive_cascade
from element_ref import EarRef, FaceRef, FeatureRef
from nubovca import synth

pytestmark = pytest.mark.gpu

STANDINS = {"haarcascade_mcs_righteye.xml": (18, 12, 0), "haarcascade_mcs_lefteye.xml": (18, 12, 1),
            "haarcascade_mcs_mouth.xml": (25, 15, 4), "haarcascade_mcs_nose.xml": (18, 15, 5),
            "haarcascade_mcs_rightear.xml": (12, 20, 0), "haarcascade_mcs_leftear.xml": (12, 20, 1)}


@pytest.fixture(scope="module")
def cdir(tmp_path_factory, cascade_dir):
    d = tmp_path_factory.mktemp("cascades")
    shutil.copy(os.path.join(cascade_dir, "haarcascade_frontalface_alt.xml"), d / "haarcascade_frontalface_alt.xml")
    # the synthetic faces are frontal: the frontal model also stands in for the profile model of the ear element
    shutil.copy(os.path.join(cascade_dir, "haarcascade_frontalface_alt.xml"), d / "haarcascade_profileface.xml")
    for name, (w, h, seed) in STANDINS.items():
        permissive_cascade(str(d / name), np.random.default_rng(seed), w, h)
    return str(d)


def oc(cdir, name):
    return O.Cascade(os.path.join(cdir, name))


def sequence(W, H, k, seed, n, **kw):
    """n frames of the same scene with a little per-frame sensor noise (drives the temporal logic)."""
    base = synth.frame(W, H, k, seed, **kw)
    rng = np.random.default_rng(seed + 77)
    out = []
    for i in range(n):
        f = base.astype(np.int16) + rng.integers(-3, 4, base.shape, dtype=np.int16)
        out.append(np.clip(f, 0, 255).astype(np.uint8))
    return out


def test_face_element_sequence(cdir):
    frames = sequence(640, 480, 4, 1, 6) + [np.full((480, 640, 3), 90, np.uint8)] * 3
    for x4 in (4, 2):
        e = nv.Element("nubofacedetector", 0, cdir)
        ref = FaceRef(oc(cdir, "haarcascade_frontalface_alt.xml"))
        e.set("process-x-every-4-frames", x4); ref.p["x4"] = x4
        seen = 0
        for i, f in enumerate(frames):
            msg, pushed, sig = e.process(f, pts_ns=i * 33_000_000)
            assert pushed and sig is None
            assert msg == ref.process(f), (x4, i)
            seen += len(msg)
        assert seen > 0
        e.close()


def test_face_element_signal_and_detect_event(cdir):
    e = nv.Element("nubofacedetector", 0, cdir)
    f = synth.frame(640, 480, 4, 1)
    e.set("activate-events", 1); e.set("events-ms", 1000)
    msg, _, sig = e.process(f, now_ms=1e15)
    assert sig == "".join(f"x:{m[2]},y:{m[3]},width:{m[4]},height:{m[5]};" for m in msg) and len(msg) >= 1
    assert e.process(f, now_ms=1e15 + 500)[2] is None                 # rate limited
    assert e.process(f, now_ms=1e15 + 1501)[2] is not None
    e.close()
    e = nv.Element("nubofacedetector", 0, cdir)
    e.set("detect-event", 1)
    assert e.process(f)[0] == []                                      # waits for an upstream "motion" event
    e.push_motion()
    assert len(e.process(f)[0]) >= 1
    e.close()


@pytest.mark.parametrize("kind,factory,files", [
    ("eye", "nuboeyedetector", ("haarcascade_mcs_righteye.xml", "haarcascade_mcs_lefteye.xml")),
    ("mouth", "nubomouthdetector", ("haarcascade_mcs_mouth.xml",)),
    ("nose", "nubonosedetector", ("haarcascade_mcs_nose.xml",))])
def test_feature_elements_cfg2(cdir, kind, factory, files):
    """BASELINE config 2: one 1280x720 stream, face stage at 160 wide, features at 320 wide inside the faces."""
    frames = sequence(1280, 720, 3, 2, 5, smin=0.4, smax=0.6)      # faces >= 30 px at the 160-wide face stage
    e = nv.Element(factory, 0, cdir)
    ref = FeatureRef(kind, oc(cdir, "haarcascade_frontalface_alt.xml"), *[oc(cdir, f) for f in files])
    total = 0
    for i, f in enumerate(frames):
        msg, pushed, _ = e.process(f, pts_ns=i * 33_000_000)
        exp = ref.process(f)
        assert pushed and msg == exp, (kind, i, msg[:4], exp[:4])
        total += sum(1 for m in msg if m[1] != "face")
    assert total > 0, "stand-in cascade never fired: the test would be vacuous"
    e.close()


def test_feature_elements_with_tree_and_tilted_models(tmp_path, cascade_dir):
    """The nested elements with REAL feature models under the reference's file names: haarcascade_smile.xml (tilted
    features) as the mouth model, righteye/lefteye_2splits (trees + tilted features) as the eye models.  These run the
    predictOrdered kernels and the tilted integral, per ROI, on the auxiliary streams."""
    d = tmp_path
    face = "haarcascade_frontalface_alt.xml"
    shutil.copy(os.path.join(cascade_dir, face), d / face)
    for src, dst in [("haarcascade_smile.xml", "haarcascade_mcs_mouth.xml"), ("haarcascade_righteye_2splits.xml", "haarcascade_mcs_righteye.xml"),
                     ("haarcascade_lefteye_2splits.xml", "haarcascade_mcs_lefteye.xml")]:
        shutil.copy(os.path.join(cascade_dir, src), d / dst)
    frames = sequence(1280, 720, 3, 2, 3, smin=0.4, smax=0.6)
    e = nv.Element("nubomouthdetector", 0, str(d))
    ref = FeatureRef("mouth", oc(str(d), face), oc(str(d), "haarcascade_mcs_mouth.xml"))
    total = 0
    for i, f in enumerate(frames):
        msg, _, _ = e.process(f)
        assert msg == ref.process(f), i
        total += sum(1 for m in msg if m[1] == "mouth")
    assert total > 0
    e.close()
    frames = sequence(1280, 720, 2, 3, 3, smin=0.5, smax=0.9)
    e = nv.Element("nuboeyedetector", 0, str(d))
    e.set("width-to-process", 640)
    ref = FeatureRef("eye", oc(str(d), face), oc(str(d), "haarcascade_mcs_righteye.xml"), oc(str(d), "haarcascade_mcs_lefteye.xml"))
    ref.p["w2p"] = 640
    total = 0
    for i, f in enumerate(frames):
        msg, _, _ = e.process(f)
        assert msg == ref.process(f), i
        total += len(msg)
    assert total > 0
    e.close()


def test_feature_element_faces_from_upstream_event(cdir):
    """ROI nesting through the downstream metadata event: the face element's rectangles (original-image
    coordinates) feed the mouth element in detect-event mode (kmseyedetect.cpp:954-961 pattern)."""
    frames = sequence(1280, 720, 3, 2, 3, smin=0.4, smax=0.6)
    face = nv.Element("nubofacedetector", 0, cdir)
    mouth = nv.Element("nubomouthdetector", 0, cdir)
    mouth.set("detect-event", 1)
    ref = FeatureRef("mouth", None, oc(cdir, "haarcascade_mcs_mouth.xml"))
    ref.p["detect_event"] = 1
    assert mouth.process(frames[0])[0] == []                          # nothing queued: frame skipped
    total = 0
    for f in frames:
        fmsg, _, _ = face.process(f)
        rects = [[m[2], m[3], m[4], m[5]] for m in fmsg if m[1] == "face"]
        mouth.push_faces(rects); ref.queue.append(rects)
        msg, _, _ = mouth.process(f)
        assert msg == ref.process(f)
        total += sum(1 for m in msg if m[1] == "mouth")
    assert total > 0
    face.close(); mouth.close()


def test_ear_element(cdir):
    frames = sequence(1280, 720, 3, 2, 3, smin=0.4, smax=0.6)
    e = nv.Element("nuboeardetector", 0, cdir)
    ref = EarRef(oc(cdir, "haarcascade_profileface.xml"), oc(cdir, "haarcascade_mcs_rightear.xml"),
                 oc(cdir, "haarcascade_mcs_leftear.xml"))
    for x in (e,):
        x.set("multi-scale-factor", 10)
    ref.p["sf"] = 10
    total = 0
    for i, f in enumerate(frames):
        msg, pushed, _ = e.process(f)
        assert not pushed                                             # built but never pushed (kmseardetect.cpp:210-290)
        assert msg == ref.process(f), i
        total += sum(1 for m in msg if m[1] == "ear")
    assert total > 0
    e.close()


def test_tracker_element(cdir):
    frames = synth.tracker_sequence(640, 360, 5, seed=5)
    e = nv.Element("nubotracker", 0, cdir)
    st = O.TrackerState(640, 360)
    e.set("activate-events", 1); e.set("events-ms", 0)
    for i, f in enumerate(frames):
        msg, pushed, sig = e.process(f, pts_ns=40_000_000 * (i + 1), now_ms=1e15 + 40.0 * i)   # MHI time = buffer PTS
        exp, _, _ = st.process(f, 40.0 * (i + 1))
        assert [list(m[2:]) for m in msg] == exp.tolist(), i
        assert (sig is not None) == (len(exp) > 0)
    e.close()


def test_view_properties_draw_into_the_frame(cdir):
    """view-faces / view-mouths / set_visual_mode: the frame comes back with the reference's cvRectangle drawing
    (checked against cv2.rectangle on the emitted rectangles), and is untouched when the property is off."""
    cv2 = pytest.importorskip("cv2")
    f = synth.frame(640, 480, 4, 1)
    e = nv.Element("nubofacedetector", 0, cdir)
    g = f.copy()
    msg, _, _ = e.process(g)
    assert (g == f).all() and len(msg) >= 1
    e.set("view-faces", 1)
    g = f.copy()
    msg, _, _ = e.process(g)
    exp = f.copy()
    for m in msg:                                  # BaseFace.cpp:76: (x, y) .. (x + w - 1, y + h - 1) in processing units x scale
        x, y, w, h = (v // 4 for v in m[2:])
        cv2.rectangle(exp, (x * 4, y * 4), ((x + w - 1) * 4, (y + h - 1) * 4), (255, 128, 0), 3, 8, 0)
    assert (g == exp).all() and (g != f).any()
    e.close()

    frames = sequence(1280, 720, 3, 2, 2, smin=0.4, smax=0.6)
    e = nv.Element("nubomouthdetector", 0, cdir)
    e.set("view-mouths", 1)
    cols = [(0, 255, 255), (0, 128, 255), (0, 0, 255), (255, 0, 255), (255, 128, 0), (255, 0, 0), (255, 255, 0), (0, 255, 0)]
    drawn = 0
    for fr in frames:
        g = fr.copy()
        msg, _, _ = e.process(g)
        exp = fr.copy()
        for j, m in enumerate([m for m in msg if m[1] == "mouth"]):
            cv2.rectangle(exp, (m[2], m[3]), (m[2] + m[4] - 1, m[3] + m[5] - 1), cols[j % 8], 3, 8, 0)
            drawn += 1
        assert (g == exp).all()
    assert drawn > 0
    e.close()

    e = nv.Element("nuboeyedetector", 0, cdir)
    e.set("view-eyes", 1)
    drawn = 0
    for fr in frames:
        g = fr.copy()
        msg, _, _ = e.process(g)
        exp = fr.copy()
        rights = [m for m in msg if m[0] == "eye_right"]; lefts = [m for m in msg if m[0] == "eye_left"]
        radius = -1                                   # kmseyedetect.cpp:1069-1101: first right eye, first left eye, one radius
        for lst in (rights, lefts):
            if lst:
                m = lst[0]
                if radius < 0:
                    radius = int(np.rint((m[4] + m[5]) * 0.25))
                cv2.circle(exp, (m[2] + m[4] // 2, m[3] + m[5] // 2), radius, (255, 0, 0), 4, 8, 0)
                drawn += 1
        assert (g == exp).all()
    assert drawn > 0
    e.close()

    seq = synth.tracker_sequence(640, 360, 4, seed=5)
    e = nv.Element("nubotracker", 0, cdir)
    e.set("set_visual_mode", 1)
    drawn = 0
    for i, fr in enumerate(seq):
        g = fr.copy()
        msg, _, _ = e.process(g, pts_ns=40_000_000 * (i + 1))
        exp = fr.copy()
        for m in msg:
            cv2.rectangle(exp, (m[2], m[3]), (m[2] + m[4], m[3] + m[5]), (0, 0, 255, 0), 3, 8, 0)
            drawn += 1
        assert (g == exp).all(), i
    assert drawn > 0
    e.close()


def test_device_resident_frames_and_device_side_overlay(cdir):
    """Frames that never leave the GPU (nv_element_transform_frame_device): every element gives the messages it gives on
    the same frame in host memory, and its view-* / set_visual_mode overlay — written by k_draw_spans into the device
    frame — equals the host drawing pixel for pixel (itself pinned to cv2 above).  Then nv_draw_shapes_device against
    cv2.rectangle / cv2.circle on random overlapping, clipped shapes."""
    torch = pytest.importorskip("torch")
    cv2 = pytest.importorskip("cv2")
    frames = sequence(1280, 720, 3, 2, 3, smin=0.4, smax=0.6)
    drawn = 0
    for factory, prop in [("nubofacedetector", "view-faces"), ("nubomouthdetector", "view-mouths"), ("nubonosedetector", "view-noses"),
                          ("nuboeyedetector", "view-eyes"), ("nuboeardetector", "view-ears")]:
        eh, ed = nv.Element(factory, 0, cdir), nv.Element(factory, 0, cdir)
        eh.set(prop, 1); ed.set(prop, 1)
        for i, fr in enumerate(frames):
            g = fr.copy()
            out_h = eh.process(g, pts_ns=i * 40_000_000, now_ms=1e15 + 40.0 * i)
            d = torch.from_numpy(fr.copy()).cuda()
            out_d = ed.process_device(d.data_ptr(), fr.shape[1], fr.shape[0], fr.strides[0], pts_ns=i * 40_000_000, now_ms=1e15 + 40.0 * i)
            assert out_d == out_h, (factory, i)
            got = d.cpu().numpy()
            assert (got == g).all(), (factory, i)
            drawn += int((g != fr).any())
        eh.close(); ed.close()
    assert drawn >= 8
    seq = synth.tracker_sequence(640, 360, 5, seed=5)
    eh, ed = nv.Element("nubotracker", 0, cdir), nv.Element("nubotracker", 0, cdir)
    eh.set("set_visual_mode", 1); ed.set("set_visual_mode", 1)
    drawn = 0
    for i, fr in enumerate(seq):
        g = fr.copy()
        out_h = eh.process(g, pts_ns=40_000_000 * (i + 1))
        d = torch.from_numpy(fr.copy()).cuda()
        out_d = ed.process_device(d.data_ptr(), fr.shape[1], fr.shape[0], fr.strides[0], pts_ns=40_000_000 * (i + 1))
        assert out_d == out_h and (d.cpu().numpy() == g).all(), i
        drawn += int((g != fr).any())
    assert drawn > 0
    eh.close(); ed.close()
    with pytest.raises(nv.NuboError):                                  # a host pointer is refused, not dereferenced on the device
        e = nv.Element("nubofacedetector", 0, cdir)
        e.process_device(frames[0].ctypes.data, 1280, 720, 3840)

    ctx = nv.Context(0, 640, 480)
    rng = np.random.default_rng(23)
    for t in range(25):
        W, H, cn = int(rng.integers(40, 640)), int(rng.integers(40, 480)), 3 + t % 2
        shapes = []
        for _ in range(int(rng.integers(1, 14))):
            col = tuple(int(v) for v in rng.integers(0, 256, 3))
            if rng.integers(0, 3):
                x0, x1 = (int(v) for v in rng.integers(-10, W + 10, 2)); y0, y1 = (int(v) for v in rng.integers(-10, H + 10, 2))
                shapes.append(("rect", x0, y0, x1, y1, col))
            else:
                shapes.append(("circle", int(rng.integers(-20, W + 20)), int(rng.integers(-20, H + 20)), int(rng.integers(0, 60)),
                               4 if rng.integers(0, 4) else int(rng.integers(2, 9)), col))
        base = rng.integers(0, 256, (H, W, cn), dtype=np.uint8)
        exp = base.copy()
        for (kind, a, b, c, dd, col) in shapes:
            sc = col + ((0,) if cn == 4 else ())
            if kind == "rect":
                cv2.rectangle(exp, (a, b), (c, dd), sc, 3, 8, 0)
            else:
                cv2.circle(exp, (a, b), c, sc, dd, 8, 0)
        d = torch.from_numpy(base.copy()).cuda()
        ctx.draw_shapes_device(d.data_ptr(), W, H, base.strides[0], cn, shapes)
        assert (d.cpu().numpy() == exp).all(), (t, shapes)
    ctx.close()


def test_elements_on_concurrent_host_threads(cdir):
    """One streaming thread per element, as GStreamer runs them: four host threads, each driving its own face and
    mouth element over its own frames at the same time (shared cascade files, separate contexts and streams), must
    see exactly what the same elements produce one after the other."""
    import threading
    nthreads = 4
    seqs = [sequence(1280, 720, 3, 2 + t, 4, smin=0.4, smax=0.6) for t in range(nthreads)]

    def run(frames):
        face = nv.Element("nubofacedetector", 0, cdir)
        mouth = nv.Element("nubomouthdetector", 0, cdir)
        out = []
        for f in frames:
            out.append((face.process(f)[0], mouth.process(f)[0]))
        face.close(); mouth.close()
        return out

    want = [run(s_) for s_ in seqs]
    got = [None] * nthreads
    errs = []

    def worker(t):
        try:
            got[t] = run(seqs[t])
        except Exception as ex:          # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    assert got == want
    assert sum(len(m) for w in want for (fm, mm) in w for m in (fm, mm)) > 0
