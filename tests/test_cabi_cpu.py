"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/nubovca.h
declares, parses cascades exactly like the oracle's independent parser, reports errors without
throwing, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nubovca as nv
import oracle as O
from cascade_xml_util import random_cascade, random_lbp_cascade, write_cascade, write_old_format

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "nubovca.h")).read()
    declared = set(re.findall(r"NV_API\s+[\w\s\*]+?\b(nv_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(nv.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in nubovca.h but not exported"
    assert declared == set(nv.EXPORTS), declared ^ set(nv.EXPORTS)
    assert nv.version().startswith("nubovca-b200")


@pytest.mark.parametrize("name", ["haarcascade_frontalface_alt.xml", "haarcascade_profileface.xml",
                                  "haarcascade_eye.xml", "haarcascade_frontalface_default.xml"])
def test_cascade_loader_matches_independent_parser(name, cascade_dir):
    path = os.path.join(cascade_dir, name)
    c = nv.Cascade(path); d = O.parse_cascade_xml(path)
    assert (c.info.win_w, c.info.win_h) == (d["win_w"], d["win_h"])
    assert c.info.nstages == len(d["stage_ntrees"]) and c.info.nstumps == len(d["stump_feat"])
    assert c.info.nfeatures == len(d["feat_rect"])
    assert c.info.n3rect == int((d["feat_weight"][:, 2] != 0).sum())
    for s in range(c.info.nstages):
        nt, thr = c.stage(s)
        assert nt == d["stage_ntrees"][s]
        assert thr == np.float32(d["stage_thr"][s]) - np.float32(1e-5)
    for i in range(c.info.nstumps):
        r, w, t = c.stump(i)
        f = d["stump_feat"][i]
        assert (r == d["feat_rect"][f]).all() and (w == d["feat_weight"][f]).all()
        assert (t == np.array([d["stump_thr"][i], d["stump_left"][i], d["stump_right"][i]], np.float32)).all()


def test_frontalface_alt_shape(cascade_dir):
    # 20x20, 22 stages, 2135 stumps, 360 three-rect features
    c = nv.Cascade(os.path.join(cascade_dir, "haarcascade_frontalface_alt.xml"))
    assert (c.info.win_w, c.info.win_h, c.info.nstages, c.info.nstumps, c.info.n3rect) == (20, 20, 22, 2135, 360)
    assert [c.stage(s)[0] for s in range(22)] == [3, 16, 21, 39, 33, 44, 50, 51, 56, 71, 80, 103, 111, 102, 135, 137,
                                                  140, 160, 177, 182, 211, 213]
    assert c.info.order_free_sums == 1


def test_order_free_certificate(tmp_path):
    p = str(tmp_path / "a.xml")
    write_cascade(p, 20, 20, [(0.5, [(0, 1e30, 1e8, 1e8), (0, 1e30, 1.0, 1.0), (0, 1e30, -1e8, -1e8)])],
                  [[(2, 2, 10, 10, -1.0), (4, 4, 5, 5, 2.0)]])
    assert nv.Cascade(p).info.order_free_sums == 1          # 1e8 + 1 is exact in double
    write_cascade(p, 20, 20, [(0.5, [(0, 1e30, 1e30, 1e30), (0, 1e30, 1.0, 1.0)])],
                  [[(2, 2, 10, 10, -1.0), (4, 4, 5, 5, 2.0)]])
    assert nv.Cascade(p).info.order_free_sums == 0          # 1e30 + 1 rounds: order matters


def test_random_cascade_roundtrip(tmp_path):
    p = str(tmp_path / "r.xml")
    random_cascade(p, np.random.default_rng(0), nstages=5, max_trees=7)
    c = nv.Cascade(p); d = O.parse_cascade_xml(p)
    assert c.info.nstumps == len(d["stump_feat"])
    for i in range(c.info.nstumps):
        r, w, t = c.stump(i)
        assert (r == d["feat_rect"][d["stump_feat"][i]]).all() and (w == d["feat_weight"][d["stump_feat"][i]]).all()


def test_old_format_cascade(tmp_path, cascade_dir):
    """The OpenCV 2.4 systems the reference ran on ship opencv-haar-classifier XML; it must load to the same model."""
    src = os.path.join(cascade_dir, "haarcascade_frontalface_alt.xml")
    d = O.parse_cascade_xml(src)
    p = str(tmp_path / "old.xml")
    write_old_format(p, d)
    new, old, od = nv.Cascade(src), nv.Cascade(p), O.parse_cascade_xml(p)
    assert (old.info.win_w, old.info.nstages, old.info.nstumps, old.info.n3rect) == (20, 22, 2135, 360)
    for s_ in range(22):
        assert new.stage(s_) == old.stage(s_)
    for i in range(0, 2135, 7):
        a, b = new.stump(i), old.stump(i)
        assert all((x == y).all() for x, y in zip(a, b))
        assert (od["feat_rect"][od["stump_feat"][i]] == a[0]).all() and od["stump_thr"][i] == a[2][0]


def test_shipped_old_format_file():
    # cv2 ships one genuine OpenCV-1.x-layout file: 64x16 window, 16 stages, 91 stumps
    import cv2
    p = os.path.join(cv2.data.haarcascades, "haarcascade_license_plate_rus_16stages.xml")
    c = nv.Cascade(p); d = O.parse_cascade_xml(p)
    assert (c.info.win_w, c.info.win_h, c.info.nstages, c.info.nstumps) == (64, 16, 16, 91)
    assert [c.stage(s_)[0] for s_ in range(16)] == d["stage_ntrees"].tolist()
    for i in range(91):
        r, w, t = c.stump(i)
        assert (r == d["feat_rect"][i]).all() and (w == d["feat_weight"][i]).all() and t[0] == d["stump_thr"][i]


def test_cascade_error_paths(tmp_path):
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(tmp_path / "missing.xml"))
    assert e.value.code == -3
    bad = tmp_path / "bad.xml"
    bad.write_text("<opencv_storage><cascade><stageType>BOOST</stageType>")
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(bad))
    assert e.value.code == -4
    bad.write_text("this is not xml")
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(bad))
    assert e.value.code == -4
    # a tree whose child index points backwards would loop for ever: refused at load
    loop = tmp_path / "loop.xml"
    loop.write_text('<opencv_storage><cascade><stageType>BOOST</stageType><featureType>HAAR</featureType><height>20</height>'
                    '<width>20</width><stages><_><stageThreshold>0.</stageThreshold><weakClassifiers><_><internalNodes>'
                    '1 0 0 0.5 1 -1 0 0.5</internalNodes><leafValues>1. 2. 3.</leafValues></_></weakClassifiers></_></stages>'
                    '<features><_><rects><_>0 0 4 4 -1.</_><_>0 0 2 4 2.</_></rects><tilted>0</tilted></_></features>'
                    '</cascade></opencv_storage>')
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(loop))
    assert e.value.code == -4
    # a tilted rect whose rotated corners leave the window
    tl = tmp_path / "tilt.xml"
    tl.write_text(loop.read_text().replace("1 0 0 0.5 1 -1 0 0.5", "0 -1 0 0.5").replace("1. 2. 3.", "1. 2.")
                  .replace("<tilted>0</tilted>", "<tilted>1</tilted>"))
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(tl))
    assert e.value.code == -4


GENERAL = ["haarcascade_lefteye_2splits.xml", "haarcascade_righteye_2splits.xml", "haarcascade_smile.xml",
           "haarcascade_eye_tree_eyeglasses.xml", "haarcascade_frontalface_alt2.xml"]


@pytest.mark.parametrize("name", GENERAL)
def test_general_cascade_loader(name, cascade_dir, tmp_path):
    """Trees of more than one node and tilted features (OpenCV's predictOrdered path), new and old XML layout."""
    path = os.path.join(cascade_dir, name)
    d = O.parse_cascade_xml(path)
    old = str(tmp_path / "old.xml")
    write_old_format(old, d)
    for p in (path, old):
        c = nv.Cascade(p)
        assert c.info.general == 1 and c.info.has_tilted == int(d["feat_tilted"].any())
        assert c.info.nstumps == len(d["tree_nnodes"]) and c.info.nnodes == len(d["node_feat"])
        assert c.info.order_free_sums == 1          # these models' stage sums are exact in any order
        n0 = l0 = 0
        for t in range(c.info.nstumps):
            nodes, thr, leaves = c.tree(t)
            nn = int(d["tree_nnodes"][t])
            assert len(nodes) == nn
            assert (nodes[:, 1] == d["node_left"][n0:n0 + nn]).all() and (nodes[:, 2] == d["node_right"][n0:n0 + nn]).all()
            assert (thr == d["node_thr"][n0:n0 + nn]).all() and (leaves == d["leaves"][l0:l0 + nn + 1]).all()
            for i in range(nn):          # the old layout stores one feature per node: compare by content
                r, w, tilted = c.feature(int(nodes[i, 0]))
                f = d["node_feat"][n0 + i]
                assert (r == d["feat_rect"][f]).all() and (w == d["feat_weight"][f]).all() and tilted == d["feat_tilted"][f]
            n0 += nn; l0 += nn + 1


@pytest.mark.parametrize("max_nodes", [1, 3])
def test_lbp_cascade_loader(tmp_path, max_nodes):
    """BOOST/LBP cascades (categorical stumps and trees): the loader against the oracle's independent parser."""
    p = str(tmp_path / "lbp.xml")
    random_lbp_cascade(p, np.random.default_rng(5 + max_nodes), nstages=5, max_trees=9, max_nodes=max_nodes)
    c = nv.Cascade(p); d = O.parse_cascade_xml(p)
    assert d["lbp"] and c.info.lbp == 1 and c.info.general == 1 and c.info.has_tilted == 0
    assert (c.info.win_w, c.info.win_h, c.info.nstages) == (d["win_w"], d["win_h"], len(d["stage_ntrees"]))
    assert c.info.nstumps == len(d["tree_nnodes"]) and c.info.nnodes == len(d["node_feat"]) and c.info.nfeatures == len(d["feat_rect"])
    for s in range(c.info.nstages):
        assert c.stage(s) == (d["stage_ntrees"][s], np.float32(d["stage_thr"][s]) - np.float32(1e-5))
    n0 = l0 = 0
    for t in range(c.info.nstumps):
        nodes, _, leaves = c.tree(t)
        nn = int(d["tree_nnodes"][t])
        assert len(nodes) == nn and (nodes[:, 0] == d["node_feat"][n0:n0 + nn]).all()
        assert (nodes[:, 1] == d["node_left"][n0:n0 + nn]).all() and (nodes[:, 2] == d["node_right"][n0:n0 + nn]).all()
        assert (leaves == d["leaves"][l0:l0 + nn + 1]).all()
        for i in range(nn):
            assert (c.subset(n0 + i) == d["node_subset"][n0 + i]).all()
        n0 += nn; l0 += nn + 1
    for f in range(c.info.nfeatures):
        assert (c.feature(f)[0] == d["feat_rect"][f]).all()
    # error paths: a cell grid that leaves the window, a category count the evaluator does not have
    txt = open(p).read()
    bad = tmp_path / "bad.xml"
    bad.write_text(txt.replace("</rect></_>", "</rect></_>\n<_><rect>20 0 2 2</rect></_>", 1))     # 20 + 3 * 2 > 24
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(bad))
    assert e.value.code == -4
    bad.write_text(txt.replace("<maxCatCount>256</maxCatCount>", "<maxCatCount>128</maxCatCount>"))
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(bad))
    assert e.value.code == -5
    bad.write_text(txt.replace("<featureType>LBP</featureType>", "<featureType>HOG</featureType>"))
    with pytest.raises(nv.NuboError) as e:
        nv.Cascade(str(bad))
    assert e.value.code == -5


@pytest.mark.skipif(nv.device_count() > 0, reason="a GPU is visible")
def test_no_cpu_fallback():
    with pytest.raises(nv.NuboError) as e:
        nv.Context()
    assert e.value.code == nv.NV_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)
