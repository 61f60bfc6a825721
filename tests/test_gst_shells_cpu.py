"""The GStreamer shells (nubomedia-vca_b200/gst/gstnubovca.cpp) built against the mock GStreamer, next to the reference's
own elements built against the same mock (oracle/_ref): what class_init declares must agree — factory names, rank, pad
templates and caps, every property (name, type, range, pspec default, flags, the boxed image-to-overlay), the signal and
its signature, the overridden virtual functions.  Compute paths need a GPU and live in tests/test_gst_shells_gpu.py."""
import os

import pytest

import refgst


def parse(text):
    d = {"property": {}, "pad": [], "signal": [], "vfuncs": None, "factory": None}
    for line in text.strip().split("\n"):
        k, *rest = line.split("|")
        if k == "property":
            # name | type | min | max | default | flags   (ids and nicks are private to a class)
            d["property"][rest[0]] = tuple(rest[1:6])
        elif k in ("pad", "signal"):
            d[k].append(tuple(rest))
        elif k in ("vfuncs", "factory"):
            d[k] = tuple(rest)
    return d


@pytest.mark.parametrize("factory", refgst.FACTORIES)
def test_shell_declares_what_the_reference_declares(factory):
    ref, sh = parse(refgst.ref().describe(factory)), parse(refgst.shell().describe(factory))
    assert sh["factory"] == ref["factory"]                     # name and GST_RANK_NONE
    assert sh["pad"] == ref["pad"]                             # src / sink, always, identical caps strings
    assert sh["property"] == ref["property"], (set(sh["property"]) ^ set(ref["property"]))
    assert sh["signal"] == ref["signal"]
    assert sh["vfuncs"] == ref["vfuncs"]


@pytest.mark.parametrize("factory", refgst.FACTORIES)
def test_shell_property_round_trip_like_the_reference(factory):
    R, S = refgst.ref(), refgst.shell()
    r, s = R.element(factory), S.element(factory)
    names = [ln.split("|")[1] for ln in R.describe(factory).split("\n") if ln.startswith("property|") and "GstStructure" not in ln]
    for n in names:
        assert r.get(n) == s.get(n), n                          # *_init() values
    for n in names:
        for v in (0, 1, 3, 160, 640, 30000, 300000, -1, 10 ** 7):
            assert r.set(n, v) == s.set(n, v), (n, v)           # GLib range validation
            assert r.get(n) == s.get(n), (n, v)
    if factory != "nubotracker":                                # image-to-overlay: boxed GstStructure, get on a fresh element gives an empty one
        buf_r, buf_s = (C.create_string_buffer(4096) for _ in range(2))
        for h, e, b in ((R, r, buf_r), (S, s, buf_s)):
            st = h.L.mh_get_structure(e.e, b"image-to-overlay")
            h.L.mh_st_to_string(st, b, 4096)
            h.L.mh_st_free(st)
        assert buf_r.value == buf_s.value == b"image_to_overlay;"
        for h, e, b in ((R, r, buf_r), (S, s, buf_s)):
            st = h.structure("image_to_overlay", [("offsetXPercent", "double", 0.1), ("offsetYPercent", "double", 0.2), ("widthPercent", "double", 1.0),
                                                  ("heightPercent", "double", 1.0), ("url", "string", "/nonexistent.png")])
            assert h.L.mh_set_structure(e.e, b"image-to-overlay", st) == 0
            h.L.mh_st_free(st)
            st = h.L.mh_get_structure(e.e, b"image-to-overlay")
            h.L.mh_st_to_string(st, b, 4096)
            h.L.mh_st_free(st)
        assert buf_r.value == buf_s.value and b"url=(string)/nonexistent.png" in buf_s.value
    r.close(); s.close()


def test_custom_events_are_forwarded_like_the_reference():
    """kmseyedetect.cpp:192-218 hands every sink event on through gst_pad_event_default; the face element keeps custom
    downstream events to itself (kmsfacedetect.cpp:251-280); ear and tracker do not override sink_event."""
    R, S = refgst.ref(), refgst.shell()
    for factory in refgst.FACTORIES:
        counts = []
        for h in (R, S):
            e = h.element(factory)
            e.send_event(h.faces_message([(1, 2, 3, 4)]))
            e.send_event(h.structure("eos-like"), downstream_custom=False)
            counts.append(h.L.mh_pushed_count(e.e))
            e.close()
        assert counts[0] == counts[1], (factory, counts)


import ctypes as C  # noqa: E402
