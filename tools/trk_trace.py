"""Per-phase clock64 timeline of the fused tracker kernel (k_trk_fused built with -DNV_TRK_TRACE into a private copy of the
library): where a block's time goes, and how long the last block works alone.  Usage on a GPU box: python tools/trk_trace.py"""
import ctypes as C
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nubomedia-vca_b200")
OUT = "/tmp/nubovca_trk_trace"
os.makedirs(OUT, exist_ok=True)
LIB = os.path.join(OUT, "libnubovca_trace.so")
if not os.path.exists(LIB):
    objs = []
    for f in ("context", "kernels_prep", "kernels_pyramid", "kernels_cascade", "kernels_group", "kernels_tracker", "elements"):
        o = os.path.join(OUT, f + ".o")
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-DNV_TRK_TRACE",
                               "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "csrc"),
                               "-c", "-o", o, os.path.join(PKG, "csrc", f + ".cu")])
        objs.append(o)
    o = os.path.join(OUT, "cascade_xml.o")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "csrc"), "-c", "-o", o, os.path.join(PKG, "csrc", "cascade_xml.cpp")])
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + [o])
os.environ["NUBOVCA_LIB"] = LIB
sys.path.insert(0, os.path.join(PKG, "python"))
import numpy as np  # noqa: E402
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402

NAMES = ["table -> smem", "point ops + labels", "unions", "flatten", "reductions", "roots / slots / publish", "edge arrival", "stitch",
         "done counter", "last block: fold / collect / order"]
for (W, H) in [(1280, 720), (640, 360)]:
    seq = synth.tracker_sequence(W, H, 8, seed=4)
    t = nv.Context(0, W, H)
    for i in range(12):
        t.tracker_process(seq[i % len(seq)], 33.3 * (i + 1))
    nt = ((W + 63) // 64) * ((H + 31) // 32)
    buf = np.zeros((nt, 16), np.int64)
    assert nv._lib.nv_debug_trk_trace(buf.ctypes.data_as(C.c_void_p), nt) == 0
    d = np.diff(buf[:, :10], axis=1)
    print(f"{W}x{H}: {nt} blocks; cycles per phase (median / max over blocks)")
    for k in range(9):
        print(f"  {NAMES[k]:28s} {int(statistics.median(d[:, k])):8d} {int(d[:, k].max()):8d}")
    # checkpoint 10 is written by the frame's last block only; the other rows hold stale values of earlier frames there
    # (clock64 is per SM, so values of different blocks do not compare): the last block is the row whose 10 follows its 9
    last = int(np.nonzero(buf[:, 10] > buf[:, 9])[0][0])
    print(f"  {NAMES[9]:28s} {int(buf[last, 10] - buf[last, 9]):8d}  (block {last})")
    print(f"  whole block, median {int(statistics.median(buf[:, 9] - buf[:, 0]))}, max {int((buf[:, 9] - buf[:, 0]).max())} cycles")
    t.close()
