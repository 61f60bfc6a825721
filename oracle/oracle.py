"""ctypes front-end of the CPU oracle (oracle/nubo_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product under nubomedia-vca_b200/.

The cascade XML is parsed here with xml.etree (an independent parser from the product's C++
loader in nubomedia-vca_b200/csrc/cascade_xml.cpp, so the two cross-check each other).
XML grammar: OpenCV "opencv-cascade-classifier" new format, the files the reference loads
with CascadeClassifier::load (kmsfacedetect.cpp:163-177).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libnubo_oracle.so")

DEPTH_VARREJ = -100
DEPTH_SKIPPED = -32768


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nubo_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


class _CCascade(C.Structure):
    _fields_ = [
        ("win_w", C.c_int), ("win_h", C.c_int), ("nstages", C.c_int), ("nstumps", C.c_int),
        ("stage_ntrees", C.POINTER(C.c_int)), ("stage_thr", C.POINTER(C.c_float)),
        ("stump_feat", C.POINTER(C.c_int)), ("stump_thr", C.POINTER(C.c_float)),
        ("stump_left", C.POINTER(C.c_float)), ("stump_right", C.POINTER(C.c_float)),
        ("nfeatures", C.c_int),
        ("feat_rect", C.POINTER(C.c_int)), ("feat_weight", C.POINTER(C.c_float)),
        ("general", C.c_int),
        ("tree_nnodes", C.POINTER(C.c_int)), ("node_feat", C.POINTER(C.c_int)), ("node_thr", C.POINTER(C.c_float)),
        ("node_left", C.POINTER(C.c_int)), ("node_right", C.POINTER(C.c_int)), ("leaves", C.POINTER(C.c_float)),
        ("feat_tilted", C.POINTER(C.c_ubyte)),
        ("lbp", C.c_int), ("node_subset", C.POINTER(C.c_int)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ora_scales.restype = C.c_int
        _lib.ora_scales.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_int]
        _lib.ora_eval_level.restype = C.c_int
        _lib.ora_integral_tilted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.ora_eval_level.argtypes = [C.POINTER(_CCascade), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.ora_row_limit.restype = C.c_int
        _lib.ora_row_limit.argtypes = [C.c_int, C.c_int, C.c_int]
        _lib.ora_feature_value.restype = C.c_int
        _lib.ora_feature_value.argtypes = [C.POINTER(_CCascade), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_void_p]
        _lib.ora_lbp_code_at.restype = C.c_int
        _lib.ora_lbp_code_at.argtypes = [C.POINTER(_CCascade), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.ora_group_rectangles.restype = C.c_int
        _lib.ora_group_rectangles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
        _lib.ora_detect_multiscale.restype = C.c_int
        _lib.ora_detect_multiscale.argtypes = [C.POINTER(_CCascade), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ora_face_process.restype = C.c_int
        _lib.ora_face_process.argtypes = [C.POINTER(_CCascade), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p]
        _lib.ora_segment_motion.restype = C.c_int
        _lib.ora_segment_motion.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_int]
        _lib.ora_join_objects.restype = C.c_int
        _lib.ora_join_objects.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_int]
        _lib.ora_tracker_process.restype = C.c_int
        _lib.ora_tracker_process.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_long, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ora_update_mhi.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
        _lib.ora_absdiff_threshold.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.ora_yuv420_to_bgr.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int, C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


# ------------------------------------------------------------------------------------------
# cascade model
# ------------------------------------------------------------------------------------------
def parse_cascade_xml(path: str) -> dict:
    """Parse a BOOST/HAAR cascade (new or old XML layout): stumps or trees, upright or tilted features.

    The dict always holds the general model (tree_nnodes, node_*, leaves, feat_tilted); `general` is False when
    every weak classifier is a stump and no feature is tilted, and then the stump_* arrays are filled too."""
    root = ET.parse(path).getroot()
    casc = root.find("cascade")
    if casc is None:
        return _parse_old_format(root)
    ftype = casc.findtext("featureType", "").strip()
    if ftype == "LBP":
        return _parse_lbp(casc)
    if ftype != "HAAR":
        raise NotImplementedError("only HAAR and LBP cascades")
    win_w, win_h = int(casc.findtext("width")), int(casc.findtext("height"))
    stage_ntrees, stage_thr, trees = [], [], []
    for st in casc.find("stages"):
        weak = st.find("weakClassifiers")
        stage_thr.append(float(st.findtext("stageThreshold")))
        n = 0
        for wc in weak:
            nodes = wc.findtext("internalNodes").split()
            leaves = [float(v) for v in wc.findtext("leafValues").split()]
            nn = len(nodes) // 4
            if len(nodes) != 4 * nn or len(leaves) != nn + 1 or nn < 1:
                raise ValueError("malformed weak classifier")
            trees.append(([(int(nodes[4 * i + 2]), float(nodes[4 * i + 3]), int(nodes[4 * i]), int(nodes[4 * i + 1]))
                           for i in range(nn)], leaves))
            n += 1
        stage_ntrees.append(n)
    rects, weights, tilted = [], [], []
    for f in casc.find("features"):
        tilted.append(1 if int((f.findtext("tilted") or "0").strip()) != 0 else 0)
        r = np.zeros((3, 4), np.int32); w = np.zeros(3, np.float32)
        for k, rc in enumerate(f.find("rects")):
            t = rc.text.split()
            r[k] = [int(t[0]), int(t[1]), int(t[2]), int(t[3])]; w[k] = np.float32(float(t[4]))
        rects.append(r); weights.append(w)
    return _model(win_w, win_h, stage_ntrees, stage_thr, trees, rects, weights, tilted)


def _parse_lbp(casc) -> dict:
    """BOOST/LBP cascade (new layout only; OpenCV 1.x had none): a node is `left right feature s0 .. s7` with the
    256-bit subset of LBP codes that go LEFT; a feature is one cell rect `x y w h` of a 3 x 3 grid of cells."""
    if int(casc.find("featureParams").findtext("maxCatCount")) != 256:
        raise NotImplementedError("LBP cascade with maxCatCount != 256")
    win_w, win_h = int(casc.findtext("width")), int(casc.findtext("height"))
    stage_ntrees, stage_thr, trees, subsets = [], [], [], []
    for st in casc.find("stages"):
        stage_thr.append(float(st.findtext("stageThreshold")))
        n = 0
        for wc in st.find("weakClassifiers"):
            nodes = wc.findtext("internalNodes").split()
            leaves = [float(v) for v in wc.findtext("leafValues").split()]
            nn = len(nodes) // 11
            if len(nodes) != 11 * nn or len(leaves) != nn + 1 or nn < 1:
                raise ValueError("malformed LBP weak classifier")
            trees.append(([(int(nodes[11 * i + 2]), 0.0, int(nodes[11 * i]), int(nodes[11 * i + 1])) for i in range(nn)], leaves))
            for i in range(nn):
                subsets.append([int(v) for v in nodes[11 * i + 3:11 * i + 11]])
            n += 1
        stage_ntrees.append(n)
    rects, weights, tilted = [], [], []
    for f in casc.find("features"):
        t = f.findtext("rect").split()
        r = np.zeros((3, 4), np.int32); r[0] = [int(t[0]), int(t[1]), int(t[2]), int(t[3])]
        rects.append(r); weights.append(np.zeros(3, np.float32)); tilted.append(0)
    return _model(win_w, win_h, stage_ntrees, stage_thr, trees, rects, weights, tilted, subsets=subsets)


def _model(win_w, win_h, stage_ntrees, stage_thr, trees, rects, weights, tilted, subsets=None) -> dict:
    """trees: list of ([(feat, thr, left, right), ...], [leaf, ...]) per weak classifier; subsets (LBP models only): eight
    int32 words per node, in node order."""
    f32 = lambda a: np.array(a, np.float64).astype(np.float32)      # noqa: E731
    general = subsets is not None or any(tilted) or any(len(nodes) != 1 for nodes, _ in trees)
    d = dict(win_w=win_w, win_h=win_h, stage_ntrees=np.array(stage_ntrees, np.int32), stage_thr=f32(stage_thr),
             feat_rect=np.ascontiguousarray(np.stack(rects)), feat_weight=np.ascontiguousarray(np.stack(weights)),
             general=bool(general),
             tree_nnodes=np.array([len(nodes) for nodes, _ in trees], np.int32),
             node_feat=np.array([n[0] for nodes, _ in trees for n in nodes], np.int32),
             node_thr=f32([n[1] for nodes, _ in trees for n in nodes]),
             node_left=np.array([n[2] for nodes, _ in trees for n in nodes], np.int32),
             node_right=np.array([n[3] for nodes, _ in trees for n in nodes], np.int32),
             leaves=f32([v for _, lv in trees for v in lv]),
             feat_tilted=np.array(tilted, np.uint8), lbp=subsets is not None,
             node_subset=(np.array(subsets, np.int64).astype(np.uint32).view(np.int32).reshape(-1, 8) if subsets is not None
                          else np.zeros((0, 8), np.int32)))
    if not general:
        if any(nodes[0][2] != 0 or nodes[0][3] != -1 for nodes, _ in trees):
            raise ValueError("stump with unexpected leaf indices")
        d.update(stump_feat=np.array([nodes[0][0] for nodes, _ in trees], np.int32),
                 stump_thr=f32([nodes[0][1] for nodes, _ in trees]),
                 stump_left=f32([lv[0] for _, lv in trees]), stump_right=f32([lv[1] for _, lv in trees]))
    else:
        z = np.zeros(len(trees), np.float32)
        d.update(stump_feat=np.zeros(len(trees), np.int32), stump_thr=z, stump_left=z.copy(), stump_right=z.copy())
    return d


def _parse_old_format(root) -> dict:
    """OpenCV 1.x/2.x "opencv-haar-classifier" layout (what OpenCV 2.4 shipped; cv2 4.13 converts it on load:
    one feature per node in file order, a <left_val>/<right_val> becomes the next leaf of its tree, a
    <left_node>/<right_node> the index of the child node)."""
    old = next((k for k in root if k.find("stages") is not None and k.find("size") is not None), None)
    if old is None:
        raise NotImplementedError("not a haar cascade")
    win_w, win_h = (int(t) for t in old.findtext("size").split())
    stage_ntrees, stage_thr, trees, rects, weights, tilted = [], [], [], [], [], []
    for st in old.find("stages"):
        stage_thr.append(float(st.findtext("stage_threshold")))
        n = 0
        for tree in st.find("trees"):
            nodes, leaves = [], []
            for node in tree:
                ft = node.find("feature")
                tilted.append(1 if int((ft.findtext("tilted") or "0").strip()) != 0 else 0)
                r = np.zeros((3, 4), np.int32); w = np.zeros(3, np.float32)
                for k, rc in enumerate(ft.find("rects")):
                    t = rc.text.split()
                    r[k] = [int(t[0]), int(t[1]), int(t[2]), int(t[3])]; w[k] = np.float32(float(t[4]))
                rects.append(r); weights.append(w)
                child = []
                for side in ("left", "right"):
                    v = node.find(side + "_val")
                    if v is not None:
                        child.append(-len(leaves)); leaves.append(float(v.text))
                    else:
                        child.append(int(node.findtext(side + "_node")))
                nodes.append((len(rects) - 1, float(node.findtext("threshold")), child[0], child[1]))
            trees.append((nodes, leaves))
            n += 1
        stage_ntrees.append(n)
    return _model(win_w, win_h, stage_ntrees, stage_thr, trees, rects, weights, tilted)


class Cascade:
    def __init__(self, path_or_dict):
        d = parse_cascade_xml(path_or_dict) if isinstance(path_or_dict, str) else path_or_dict
        self.d = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
        d = self.d
        self.win_w, self.win_h = d["win_w"], d["win_h"]
        self.nstages, self.nstumps = len(d["stage_ntrees"]), len(d["stump_feat"])
        self.general = bool(d["general"])
        self.lbp = bool(d.get("lbp", False))
        if "node_subset" not in d:
            d["node_subset"] = np.zeros((0, 8), np.int32)
        self.has_tilted = bool(d["feat_tilted"].any())
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        self.c = _CCascade(
            d["win_w"], d["win_h"], self.nstages, self.nstumps,
            d["stage_ntrees"].ctypes.data_as(ip), d["stage_thr"].ctypes.data_as(fp),
            d["stump_feat"].ctypes.data_as(ip), d["stump_thr"].ctypes.data_as(fp),
            d["stump_left"].ctypes.data_as(fp), d["stump_right"].ctypes.data_as(fp),
            len(d["feat_rect"]), d["feat_rect"].ctypes.data_as(ip), d["feat_weight"].ctypes.data_as(fp),
            int(d["general"]), d["tree_nnodes"].ctypes.data_as(ip), d["node_feat"].ctypes.data_as(ip),
            d["node_thr"].ctypes.data_as(fp), d["node_left"].ctypes.data_as(ip), d["node_right"].ctypes.data_as(ip),
            d["leaves"].ctypes.data_as(fp), d["feat_tilted"].ctypes.data_as(C.POINTER(C.c_ubyte)),
            int(self.lbp), d["node_subset"].ctypes.data_as(ip))


# ------------------------------------------------------------------------------------------
# image ops
# ------------------------------------------------------------------------------------------
def bgr2gray(img):
    img = _u8(img); h, w, cn = img.shape
    out = np.empty((h, w), np.uint8)
    lib().ora_bgr2gray(_p(img), w, h, img.strides[0], cn, _p(out), w)
    return out


YUV_FORMATS = {"I420": 0, "YV12": 0, "NV12": 1, "NV21": 2}


def yuv420_planes(buf, w, h, fmt="I420"):
    """Views of the planes of a packed 4:2:0 buffer of w*h*3/2 bytes (what cv2.cvtColor takes as a (h*3/2, w) image):
    (y, u, v) for I420 / YV12 (u, v in I420 meaning), (y, uv, None) for NV12 / NV21."""
    flat = _u8(buf).reshape(-1)
    assert w % 2 == 0 and h % 2 == 0 and flat.size == w * h * 3 // 2
    y = flat[:w * h].reshape(h, w)
    if fmt in ("NV12", "NV21"):
        return y, flat[w * h:].reshape(h // 2, w), None
    a = flat[w * h:w * h + w * h // 4].reshape(h // 2, w // 2)
    b = flat[w * h + w * h // 4:].reshape(h // 2, w // 2)
    return (y, a, b) if fmt == "I420" else (y, b, a)


def yuv420_to_bgr(y, u, v=None, fmt="I420"):
    """cv2.cvtColor(.., COLOR_YUV2BGR_I420 / _YV12 / _NV12 / _NV21) on explicit planes (any row strides)."""
    h, w = y.shape
    assert y.strides[1] == 1 and u.strides[1] == 1 and (v is None or v.strides[1] == 1)
    out = np.empty((h, w, 3), np.uint8)
    lib().ora_yuv420_to_bgr(YUV_FORMATS[fmt], _p(y), y.strides[0], _p(u), u.strides[0],
                            _p(v) if v is not None else None, v.strides[0] if v is not None else 0, w, h, _p(out), 3 * w)
    return out


def resize_linear(img, dw, dh):
    img = _u8(img)
    h, w = img.shape[:2]; cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((dh, dw) if img.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().ora_resize_linear(_p(img), w, h, img.strides[0], cn, _p(out), dw, dh, dw * cn)
    return out


def equalize_hist(img):
    img = _u8(img); h, w = img.shape
    out = np.empty_like(img)
    lib().ora_equalize_hist(_p(img), w, h, img.strides[0], _p(out), w)
    return out


def resize_linear_exact(img, dw, dh):
    img = _u8(img); h, w = img.shape
    out = np.empty((dh, dw), np.uint8)
    lib().ora_resize_linear_exact(_p(img), w, h, img.strides[0], _p(out), dw, dh, dw)
    return out


def integral(img):
    img = _u8(img); h, w = img.shape
    s = np.empty((h + 1, w + 1), np.int32); q = np.empty((h + 1, w + 1), np.uint32)
    lib().ora_integral(_p(img), w, h, img.strides[0], _p(s), _p(q))
    return s, q


def integral_tilted(img):
    img = _u8(img); h, w = img.shape
    t = np.empty((h + 1, w + 1), np.int32)
    lib().ora_integral_tilted(_p(img), w, h, img.strides[0], _p(t))
    return t


# ------------------------------------------------------------------------------------------
# cascade
# ------------------------------------------------------------------------------------------
def scales(W, H, casc: Cascade, scale_factor, min_size=(0, 0), max_size=(0, 0)):
    buf = np.empty(256, np.float32)
    n = lib().ora_scales(W, H, casc.win_w, casc.win_h, float(scale_factor), min_size[0], min_size[1],
                         max_size[0], max_size[1], _p(buf), 256)
    return buf[:min(n, 256)].copy()


def level_size(W, H, sc):
    lw, lh = C.c_int(), C.c_int()
    lib().ora_level_size(W, H, C.c_float(sc), C.byref(lw), C.byref(lh))
    return lw.value, lh.value


def eval_pyramid(gray, casc: Cascade, scale_factor, min_size=(0, 0), max_size=(0, 0), keep_integrals=False):
    """Per-level artefacts the GPU path must match bit-exactly.

    Returns a list of dicts: scale, ystep, lw, lh, image, (sum, sqsum), depth [ny, nx] int16,
    cand [k,4] (un-clipped candidate rects in y -> x order).
    """
    gray = _u8(gray); H, W = gray.shape
    out = []
    scs = scales(W, H, casc, scale_factor, min_size, max_size)
    nstripes = 1
    if len(scs):
        nstripes = (max(level_size(W, H, scs[0])[0] + 1 - casc.win_w, 0) + 31) // 32
    for sc in scs:
        lw, lh = level_size(W, H, sc)
        rx, ry = lw + 1 - casc.win_w, lh + 1 - casc.win_h
        if rx <= 0 or ry <= 0:
            continue
        img = gray.copy() if (lw, lh) == (W, H) else resize_linear_exact(gray, lw, lh)
        s, q = integral(img)
        t = integral_tilted(img) if casc.has_tilted else None
        ystep = 1 if sc >= 2 else 2
        ry = lib().ora_row_limit(ry, ystep, nstripes)
        nx, ny = (rx + ystep - 1) // ystep, (ry + ystep - 1) // ystep
        depth = np.empty((ny, nx), np.int16)
        cand = np.empty((nx * ny, 4), np.int32)
        nc = C.c_int(0)
        lib().ora_eval_level(C.byref(casc.c), _p(s), _p(q), _p(t) if t is not None else None, lw, lh, ystep, C.c_float(sc), nstripes, _p(depth),
                             _p(cand), nx * ny, C.byref(nc))
        lv = dict(scale=float(sc), ystep=ystep, lw=lw, lh=lh, image=img, depth=depth, cand=cand[:nc.value].copy())
        if keep_integrals:
            lv["sum"], lv["sqsum"], lv["tilted"] = s, q, t
        out.append(lv)
    return out


def feature_value(casc: Cascade, s, q, x, y, f):
    """Normalised value of feature f at window (x, y) of a level, or None if variance-rejected."""
    out = C.c_float(0)
    ok = lib().ora_feature_value(C.byref(casc.c), _p(s), _p(q), s.shape[1], x, y, f, C.byref(out))
    return np.float32(out.value) if ok else None


def lbp_code(casc: Cascade, s, x, y, f):
    """8-bit LBP code of feature f at window (x, y) of a level with integral s (LBPEvaluator::OptFeature::calc)."""
    return int(lib().ora_lbp_code_at(C.byref(casc.c), _p(s), s.shape[1], x, y, f))


def group_rectangles(rects, thr, eps=0.2):
    r = np.ascontiguousarray(np.asarray(rects, np.int32).reshape(-1, 4)).copy()
    w = np.zeros(max(len(r), 1), np.int32)
    n = lib().ora_group_rectangles(_p(r), len(r), thr, eps, _p(w))
    return r[:n].copy(), w[:n].copy()


def detect_multiscale(gray, casc: Cascade, scale_factor=1.1, min_neighbors=3, min_size=(0, 0), max_size=(0, 0),
                      cap=1 << 20, return_nwindows=False):
    gray = _u8(gray); H, W = gray.shape
    out = np.empty((cap, 4), np.int32)
    nwin = C.c_longlong(0)
    n = lib().ora_detect_multiscale(C.byref(casc.c), _p(gray), W, H, gray.strides[0], float(scale_factor),
                                    min_neighbors, min_size[0], min_size[1], max_size[0], max_size[1],
                                    _p(out), cap, None, C.byref(nwin))
    res = out[:min(n, cap)].copy()
    return (res, nwin.value) if return_nwindows else res


def face_process(bgr, casc: Cascade, width_to_process=160, scale_factor=1.25, min_neighbors=3, min_size=None,
                 cap=4096):
    """kmsfacedetect.cpp:770-811 end to end.  Returns (rects, equalised gray at processing size)."""
    bgr = _u8(bgr); H, W, _ = bgr.shape
    iscale = W // width_to_process
    sc = float(iscale) if iscale > 0 else 1.0
    rows = int(np.rint(H / sc)) or H; cols = int(np.rint(W / sc)) or W
    geq = np.empty((rows, cols), np.uint8)
    out = np.empty((cap, 4), np.int32)
    mw, mh = (-1, -1) if min_size is None else min_size
    n = lib().ora_face_process(C.byref(casc.c), _p(bgr), W, H, bgr.strides[0], width_to_process,
                               float(scale_factor), min_neighbors, mw, mh, _p(out), cap, _p(geq))
    return out[:n].copy(), geq


# ------------------------------------------------------------------------------------------
# tracker
# ------------------------------------------------------------------------------------------
def segment_motion(mhi, ts, seg_thresh=32.0, cap=65536):
    mhi = np.ascontiguousarray(mhi, np.float32); h, w = mhi.shape
    labels = np.empty((h, w), np.int32); rects = np.empty((cap, 4), np.int32)
    n = lib().ora_segment_motion(_p(mhi), w, h, float(ts), float(seg_thresh), _p(labels), _p(rects), cap)
    return rects[:min(n, cap)].copy(), labels


def join_objects(rects, min_area, max_area, distance):
    r = np.ascontiguousarray(np.asarray(rects, np.int32).reshape(-1, 4)).copy()
    n = lib().ora_join_objects(_p(r), len(r), min_area, max_area, distance)
    return r[:n].copy()


class TrackerState:
    """Per-element temporal state of gst_nubo_tracker_process (gstnubotracker.cpp:339-421)."""

    def __init__(self, w, h):
        self.w, self.h = w, h
        self.prev = np.zeros((h, w), np.uint8)
        self.mhi = np.zeros((h, w), np.float32)
        self.num_frames = 0

    def process(self, bgra, ts, threshold=20, min_area=50, max_area=30000, distance=35, cap=65536):
        bgra = _u8(bgra)
        rects = np.empty((cap, 4), np.int32); nraw = C.c_int(0)
        mask = np.zeros((self.h, self.w), np.uint8)
        n = lib().ora_tracker_process(_p(bgra), self.w, self.h, bgra.strides[0], int(self.num_frames == 0),
                                      _p(self.prev), _p(self.mhi), float(ts), threshold, min_area, max_area,
                                      distance, _p(rects), cap, C.byref(nraw), _p(mask))
        self.num_frames += 1
        return rects[:n].copy(), nraw.value, mask
