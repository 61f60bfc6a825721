"""Latency of the generic (tree / tilted) cascade path next to cv2 on the same image (one stream, host image in)."""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nubovca as nv
import oracle as O
from nubovca import synth
import cv2
cv2.setNumThreads(os.cpu_count() or 1)
c = nv.Context(0, 1920, 1080)
cd = os.path.join(ROOT, "nubomedia-vca_b200", "cascades")
for (W, H) in [(320, 180), (640, 480), (1920, 1080)]:
    g = O.equalize_hist(O.bgr2gray(synth.frame(W, H, 4, 9, smin=0.2, smax=0.5)))
    for name in ["haarcascade_frontalface_alt.xml", "haarcascade_frontalface_alt2.xml", "haarcascade_smile.xml", "haarcascade_lefteye_2splits.xml"]:
        nc = nv.Cascade(os.path.join(cd, name)); cc = cv2.CascadeClassifier(os.path.join(cd, name))
        for _ in range(3): c.detect_multiscale(nc, g, 1.1, 3)
        n = 20 if W < 1000 else 5
        t = time.perf_counter()
        for _ in range(n): r = c.detect_multiscale(nc, g, 1.1, 3)
        tg = (time.perf_counter() - t) / n
        for _ in range(2): cc.detectMultiScale(g, scaleFactor=1.1, minNeighbors=3)
        t = time.perf_counter()
        for _ in range(n): cc.detectMultiScale(g, scaleFactor=1.1, minNeighbors=3)
        tc = (time.perf_counter() - t) / n
        print(f"{W}x{H} {name:40s} gpu {tg*1e3:8.3f} ms  cv2({cv2.getNumThreads()} thr) {tc*1e3:8.3f} ms  x{tc/tg:6.1f}  ({len(r)} rects)")
