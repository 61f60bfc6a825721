"""Per-phase clock64 timeline of full-depth windows in k_cascade_tail_fast (the warp-per-window kernel).  Needs a build of
kernels_cascade.cu with -DNV_TAIL_TRACE (the instrumentation compiles out otherwise) linked into a variant library:

    cd nubomedia-vca_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false \\
        -Xcompiler -fPIC,-fvisibility=hidden -I../include -Icsrc -DNV_TAIL_TRACE -c -o build/kc_trace.o csrc/kernels_cascade.cu
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libnubovca_trace.so build/kc_trace.o \\
        $(ls build/*.o | grep -v kernels_cascade.o | grep -v kc_trace.o)
    NUBOVCA_LIB=$PWD/lib/libnubovca_trace.so python ../tools/tail_trace.py

Prints, for the windows that pass every stage, mean and maximum cycles per phase: queue fetch + patch copy, the wait for a
round's classifier records, shared-memory reads + integer feature arithmetic, the double-precision accumulate, the stage-end
shuffle reduction and the threshold test (results of the round: profiles/r1_v5_summary.md)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import numpy as np
import nubovca as nv
from nubovca import synth
lib = C.CDLL(nv.LIB_PATH)
casc = nv.Cascade(os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml"))
ctx = nv.Context(0, 1920, 1080); ctx.set_profile(True)
buf = (C.c_longlong * (64 * 16))()
names = ["total", "to_copied", "stage_top(meta first)", "wait_rec", "lds+int", "fp64_acc", "reduce", "thr(meta)", "rounds", "stages"]
for name, (f, w2p, sf, ms) in {"noface1080": (synth.frame(1920, 1080, 0, 3), 1920, 1.1, (24, 24)), "cfg1": (synth.frame(640, 480, 4, 1), 160, 1.25, None)}.items():
    for _ in range(3):
        lib.nv_debug_tail_trace(buf, 64)
        r = ctx.face_detect(casc, f, w2p, sf, 3, ms)
    n = lib.nv_debug_tail_trace(buf, 64)
    a = np.frombuffer(buf, dtype=np.int64).reshape(64, 16)[:min(n, 64), :10].copy()
    print(name, "traced", n, "tail stage us", round(1e3 * ctx.stage_times()["cascade_tail"], 1))
    print("   mean:", {k: int(v) for k, v in zip(names, a.mean(0))})
    print("   max :", {k: int(v) for k, v in zip(names, a.max(0))})
