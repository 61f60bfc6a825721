"""tools/streams_bench.cpp: the C ABI driven from native host threads (several contexts per thread, one frame in flight
per stream).  CPU: the binary is built and refuses bad arguments.  GPU: rectangles found on BGR and on 4:2:0 frames equal
the oracle's for every (stream, frame) pair, i.e. concurrent contexts on several threads do not disturb each other."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from nubovca import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "nubomedia-vca_b200", "lib", "streams_bench")
XML = os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml")


def test_tool_is_built_and_checks_arguments():
    assert os.access(TOOL, os.X_OK), "run __graft_entry__.build()"
    r = subprocess.run([TOOL, "--fmt", "bgr"], capture_output=True, text=True)
    assert r.returncode == 2 and "bad arguments" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["bgr", "nv12", "i420"])
def test_native_threads_match_oracle(tmp_path, fmt):
    w, h, k, streams, threads, iters = 640, 360, 3, 12, 3, 7
    ocasc = O.Cascade(XML)
    frames = [synth.frame(w, h, 2, 40 + i, smin=0.3, smax=0.6) for i in range(k)]
    if fmt == "bgr":
        blobs = frames
        counts = [len(O.face_process(f, ocasc, 320, 1.25, 3, None)[0]) for f in frames]
    else:
        name = fmt.upper()
        blobs = [synth.to_yuv420(f, name) for f in frames]
        counts = [len(O.face_process(O.yuv420_to_bgr(*O.yuv420_planes(b, w, h, name), fmt=name), ocasc, 320, 1.25, 3, None)[0])
                  for b in blobs]
    assert sum(counts) > 0
    path = str(tmp_path / "frames.raw")
    with open(path, "wb") as fh:
        for b in blobs:
            fh.write(np.ascontiguousarray(b).tobytes())
    r = subprocess.run([TOOL, "--frames-file", path, "--nframes", str(k), "--fmt", fmt, "--xml", XML, "--width", str(w),
                        "--height", str(h), "--width-to-process", "320", "--streams", str(streams), "--threads", str(threads),
                        "--iters", str(iters), "--warmup", "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["frames"] == streams * iters
    assert j["rects"] == sum(counts[(s + it) % k] for s in range(streams) for it in range(iters))
