#!/usr/bin/env python3
"""Builds oracle/_ref/libnubo_ref_elements.so: the REFERENCE's own element sources — kmsfacedetect.cpp, Faces.cpp,
BaseFace.cpp, kmseyedetect.cpp, kmsmouthdetect.cpp, kmsnosedetect.cpp, kmseardetect.cpp, gstnubotracker.cpp — compiled
UNMODIFIED from where they lie under /root/reference (each is #included by path from a wrapper under oracle/refbuild/,
or handed to g++ directly; nothing is copied into this repository), against

  * tests/mock_gst/   a functional stand-in for the GLib / GObject / GStreamer slice they use,
  * oracle/refbuild/opencv2/opencv.hpp   a stand-in for the OpenCV 2.x API they call, forwarding every pixel operation
    to the CPU oracle (oracle/nubo_oracle.c, pinned to cv2 4.13) and recording the drawing calls.

The result is the reference's complete per-frame element logic (gating, ROI arithmetic, Faces::track_faces, eye / mouth /
nose merging, tracker join, event and signal payloads) as a CPU library the tests drive through tests/mock_gst/harness.cpp.
The reference's own build system is not run.  TEST INFRASTRUCTURE ONLY — the product never links it.

/root/reference does not exist on the GPU box: the prebuilt .so travels there (oracle/_ref/ is git-ignored, not
gpurun-ignored).  Without /root/reference and without a prebuilt library this script fails loudly."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NUBO_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
LIB = os.path.join(OUT, "libnubo_ref_elements.so")
MODS = {
    "face": "modules/nubo_face/nubo-face-detector/src/gst-plugins",
    "eye": "modules/nubo_eye/nubo-eye-detector/src/gst-plugins",
    "mouth": "modules/nubo_mouth/nubo-mouth-detector/src/gst-plugins",
    "nose": "modules/nubo_nose/nubo-nose-detector/src/gst-plugins",
    "ear": "modules/nubo_ear/nubo-ear-detector/src/gst-plugins",
    "tracker": "modules/nubo_tracker/nubo-tracker/src/gst-plugins",
}


def sources():
    rb = os.path.join(ROOT, "oracle", "refbuild")
    mg = os.path.join(ROOT, "tests", "mock_gst")
    own = [os.path.join(rb, f"wrap_{m}.cpp") for m in MODS] + [os.path.join(rb, "wrap_faces.cpp")]
    own += [os.path.join(mg, "minigst.cpp"), os.path.join(mg, "harness.cpp")]
    face_dir = os.path.join(REF, MODS["face"])
    direct = [os.path.join(face_dir, "Faces.cpp"), os.path.join(face_dir, "BaseFace.cpp")]
    return own, direct


def up_to_date():
    if not os.path.exists(LIB):
        return False
    own, direct = sources()
    deps = own + [os.path.join(ROOT, "tests", "mock_gst", f) for f in ("minigst.h", "prelude.h")]
    deps += [os.path.join(ROOT, "oracle", "refbuild", "opencv2", "opencv.hpp"), os.path.join(ROOT, "oracle", "refbuild", "ref_wrap.h"),
             os.path.join(ROOT, "oracle", "nubo_oracle.c"), os.path.abspath(__file__)]
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def build(verbose=False):
    if not os.path.isdir(REF):
        if os.path.exists(LIB):
            return LIB                       # GPU box: use the library built where the reference was present
        raise RuntimeError(f"{REF} is absent and {LIB} was not prebuilt: cannot build the reference elements")
    if up_to_date():
        return LIB
    os.makedirs(OUT, exist_ok=True)
    obj = os.path.join(OUT, "obj")
    os.makedirs(obj, exist_ok=True)
    own, direct = sources()
    inc = ["-I" + os.path.join(ROOT, "tests", "mock_gst"), "-I" + os.path.join(ROOT, "oracle", "refbuild")]
    inc += ["-I" + os.path.join(REF, d) for d in MODS.values()]
    cxx = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-fvisibility=hidden", "-w", "-ffp-contract=off", "-DMH_HAVE_REFCV",
           "-include", os.path.join(ROOT, "tests", "mock_gst", "prelude.h")] + inc
    objs = []
    jobs = []
    for src in own + direct:
        o = os.path.join(obj, os.path.basename(src).replace(".cpp", ".o"))
        objs.append(o)
        jobs.append((src, subprocess.Popen(cxx + ["-c", src, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    oc = os.path.join(obj, "nubo_oracle.o")
    jobs.append(("nubo_oracle.c", subprocess.Popen(
        ["gcc", "-O2", "-std=c99", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wno-misleading-indentation",
         "-c", os.path.join(ROOT, "oracle", "nubo_oracle.c"), "-o", oc], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs.append(oc)
    failed = False
    for src, p in jobs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- {src}\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("reference element build failed")
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + ["-lm"])
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
