"""CPU-side checks of the element mirrors: factory names, GObject property surface (names, ranges,
effective defaults and the reference's quirks) and the host-side face tracking."""
import numpy as np
import pytest

import nubovca as nv
from element_ref import track_faces as ref_track_faces

COMMON = {"detect-event": (0, 1, 0), "width-to-process": (0, 640, 320), "process-x-every-4-frames": (0, 4, 4),
          "multi-scale-factor": (0, 51, 25), "activate-events": (0, 1, 0), "events-ms": (0, 30000, 30001)}
SURFACE = {
    "nubofacedetector": dict(COMMON, **{"view-faces": (0, 1, 0), "send-meta-data": (0, 1, 0), "width-to-process": (0, 640, 160),
                                         "euclidean-distance": (0, 20, 8), "track-threshold": (0, 100, 40),
                                         "area-threshold": (0, 1000, 500)}),
    "nuboeyedetector": dict(COMMON, **{"view-eyes": (0, 1, 0), "send-meta-data": (0, 1, 0)}),
    "nubomouthdetector": dict(COMMON, **{"view-mouths": (0, 1, 0), "send-meta-data": (0, 1, 0)}),
    "nubonosedetector": dict(COMMON, **{"view-noses": (0, 1, 0), "send-meta-data": (0, 1, 0)}),
    # kmseardetect.cpp:1006 (the name), :943 view_ears starts at -1, events_ms is never initialised (0): values confirmed by
    # the reference's own element (tests/test_ref_elements_cpu.py)
    "nuboeardetector": dict(COMMON, **{"view-ears": (0, 1, -1), "meta-data": (0, 1, 0), "events-ms": (0, 30000, 0)}),
    "nubotracker": {"set_threshold": (0, 255, 20), "set_min_area": (0, 10000, 50), "set_max_area": (0, 300000, 30000),
                    "set_distance": (0, 2000, 35), "set_visual_mode": (0, 4, 0), "activate-events": (0, 1, 0),
                    "events-ms": (0, 30000, 30001)},
}


@pytest.mark.parametrize("factory", sorted(SURFACE))
def test_property_surface(factory, tmp_path):
    e = nv.Element(factory, 0, str(tmp_path))          # cascades missing: non-fatal, like the reference
    for name, (lo, hi, default) in SURFACE[factory].items():
        assert e.get(name) == default, name
        for bad in (lo - 1, hi + 1):
            with pytest.raises(nv.NuboError):
                e.set(name, bad)
        assert e.get(name) == default
        e.set(name, hi)
        if not (factory == "nubofacedetector" and name == "track-threshold"):
            assert e.get(name) == hi
    with pytest.raises(nv.NuboError):
        e.get("no-such-property")
    table = {n: (lo, hi, de) for n, lo, hi, de in e.properties()}          # what a shell installs its GObject properties from
    assert table == {n: (lo, hi, (de if de <= hi else de)) for n, (lo, hi, de) in SURFACE[factory].items()}
    if factory == "nuboeardetector":
        with pytest.raises(nv.NuboError):
            e.set("send-meta-data", 1)                 # the server-side name does not exist on the element
    e.close()


def test_track_threshold_setter_quirk(tmp_path):
    # kmsfacedetect.cpp:548-550: the track-threshold setter writes euclidean_threshold
    e = nv.Element("nubofacedetector", 0, str(tmp_path))
    e.set("track-threshold", 17)
    assert e.get("track-threshold") == 40 and e.get("euclidean-distance") == 17
    e.close()


def test_unknown_factory():
    with pytest.raises(nv.NuboError):
        nv.Element("nubosomethingelse")


def test_track_faces_cases():
    # same place, same size: the old rectangle is kept (no jitter), id preserved
    r, ids, nid = nv.track_faces([[100, 100, 50, 50]], [0], 1, [[101, 101, 50, 50]])
    assert r.tolist() == [[100, 100, 50, 50]] and ids.tolist() == [0] and nid == 1
    # moved beyond the size-dependent limit (area 2500 -> 3 px): the new rectangle replaces it
    r, ids, nid = nv.track_faces([[100, 100, 50, 50]], [0], 1, [[110, 100, 50, 50]])
    assert r.tolist() == [[110, 100, 50, 50]] and ids.tolist() == [0]
    # same centre, area changed by more than 15 %: old position, new size
    r, ids, nid = nv.track_faces([[100, 100, 50, 50]], [0], 1, [[95, 95, 60, 60]])
    assert r.tolist() == [[100, 100, 60, 60]] and ids.tolist() == [0]
    # farther than track-threshold: the old face is dropped, the new one gets a fresh id
    r, ids, nid = nv.track_faces([[100, 100, 50, 50]], [0], 1, [[300, 100, 50, 50]])
    assert r.tolist() == [[300, 100, 50, 50]] and ids.tolist() == [1] and nid == 2


def test_track_faces_random_against_restatement():
    rng = np.random.default_rng(0)
    for _ in range(300):
        nprev, ncur = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        prev = [[int(rng.integers(0, 200)), int(rng.integers(0, 200)), int(rng.integers(10, 90)), int(rng.integers(10, 90))]
                for _ in range(nprev)]
        cur = [[p[0] + int(rng.integers(-12, 13)), p[1] + int(rng.integers(-12, 13)), p[2] + int(rng.integers(-8, 9)),
                p[3] + int(rng.integers(-8, 9))] for p in prev[:ncur]]
        cur += [[int(rng.integers(0, 200)), int(rng.integers(0, 200)), int(rng.integers(10, 90)), int(rng.integers(10, 90))]
                for _ in range(ncur - len(cur))]
        ids = list(range(3, 3 + nprev))
        thr = int(rng.integers(5, 60))
        r, oid, nid = nv.track_faces(prev, ids, 10, cur, thr)
        exp, enid = ref_track_faces(list(zip(prev, ids)), 10, cur, thr)
        assert r.tolist() == [list(x[0]) for x in exp] and oid.tolist() == [x[1] for x in exp] and nid == enid


def test_view_rectangles_match_cv2():
    """The view-* drawing (cvRectangle, thickness 3, 8-connected) pixel by pixel against cv2.rectangle: inside,
    clipped by every border, degenerate, reversed corners, BGR and BGRA frames."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    cases = [(6, 5, 22, 16), (-3, -2, 5, 4), (4, 4, 4, 4), (30, 20, 10, 8), (0, 0, 63, 47), (60, 44, 70, 50),
             (-10, 10, 80, 12), (5, -20, 7, 80), (-9, -9, -4, -4), (10, 10, 11, 10), (10, 10, 10, 13)]
    cases += [tuple(int(v) for v in rng.integers(-8, 72, 4)) for _ in range(60)]
    for cn in (3, 4):
        for (x0, y0, x1, y1) in cases:
            base = rng.integers(0, 256, (48, 64, cn), dtype=np.uint8)
            col = tuple(int(v) for v in rng.integers(0, 256, 3))
            exp = base.copy()
            cv2.rectangle(exp, (x0, y0), (x1, y1), col + ((0,) if cn == 4 else ()), 3, 8, 0)
            got = nv.draw_rectangle(base.copy(), x0, y0, x1, y1, col)
            assert (got == exp).all(), (cn, x0, y0, x1, y1)
    with pytest.raises(nv.NuboError):
        nv.draw_rectangle(np.zeros((4, 4, 2), np.uint8), 0, 0, 1, 1, (1, 2, 3))


def test_view_eyes_circle_matches_cv2():
    """cv::circle(.., thickness 4, LINE_8) — the view-eyes drawing — pixel by pixel against cv2.circle: every radius up to
    150 at an interior centre, then random radii and centres clipped by every border, BGR and BGRA, and the other
    thicknesses the rasteriser supports."""
    cv2 = pytest.importorskip("cv2")
    for r in list(range(0, 151)) + [195, 240, 263, 274, 287, 313, 322, 326, 353, 364, 365, 367, 375, 386, 391]:
        S = 2 * r + 40
        exp = np.zeros((S, S, 3), np.uint8)
        cv2.circle(exp, (S // 2, S // 2 + 1), r, (255, 0, 0), 4, 8, 0)
        got = nv.draw_circle(np.zeros((S, S, 3), np.uint8), S // 2, S // 2 + 1, r, (255, 0, 0))
        assert (got == exp).all(), r
    rng = np.random.default_rng(8)
    for t in range(500):
        W, H, r = int(rng.integers(20, 160)), int(rng.integers(20, 160)), int(rng.integers(0, 90))
        cx, cy = int(rng.integers(-40, W + 40)), int(rng.integers(-40, H + 40))
        cn = 3 + t % 2
        th = 4 if t % 5 else int(rng.integers(2, 9))
        base = rng.integers(0, 256, (H, W, cn), dtype=np.uint8)
        col = tuple(int(v) for v in rng.integers(0, 256, 3))
        exp = base.copy()
        cv2.circle(exp, (cx, cy), r, col + ((0,) if cn == 4 else ()), th, 8, 0)
        got = nv.draw_circle(base.copy(), cx, cy, r, col, th)
        assert (got == exp).all(), (W, H, r, cx, cy, cn, th)


def _random_shapes(rng, W, H, n):
    shapes = []
    for _ in range(n):
        col = tuple(int(v) for v in rng.integers(0, 256, 3))
        if rng.integers(0, 3):
            x0, x1 = (int(v) for v in rng.integers(-10, W + 10, 2)); y0, y1 = (int(v) for v in rng.integers(-10, H + 10, 2))
            shapes.append(("rect", x0, y0, x1, y1, col))
        else:
            shapes.append(("circle", int(rng.integers(-20, W + 20)), int(rng.integers(-20, H + 20)), int(rng.integers(0, 40)),
                           4 if rng.integers(0, 4) else int(rng.integers(2, 9)), col))
    return shapes


def test_overlay_spans_match_sequential_drawing():
    """The host half of the device-side overlay (nv_draw_shapes_device, nv_element_transform_frame_device): shapes are
    rasterised into spans, the spans made disjoint with the LATER shape winning, then written in arbitrary order — the
    picture must be the one cv2 draws shape after shape, overlaps in different colours and clipping included."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(17)
    for t in range(60):
        W, H, cn = int(rng.integers(30, 200)), int(rng.integers(30, 150)), 3 + t % 2
        shapes = _random_shapes(rng, W, H, int(rng.integers(1, 12)))
        base = rng.integers(0, 256, (H, W, cn), dtype=np.uint8)
        exp = base.copy()
        for (kind, a, b, c, d, col) in shapes:
            sc = col + ((0,) if cn == 4 else ())
            if kind == "rect":
                cv2.rectangle(exp, (a, b), (c, d), sc, 3, 8, 0)
            else:
                cv2.circle(exp, (a, b), c, sc, d, 8, 0)
        got = base.copy()
        n = nv.draw_shapes_spans(got, shapes)
        assert (got == exp).all(), (t, shapes)
        assert n > 0 or (got == base).all()
    assert nv.draw_shapes_spans(np.zeros((8, 8, 3), np.uint8), []) == 0
