"""Where a small blocking call spends its time (config 1: 640x480 -> 160x120, scale 1.25): the whole call from pinned host
memory, the same with the frame already on the device (no H2D), the per-stage CUDA-event times of plain launches, and the
floor of an empty graph launch + synchronize.  Prints microseconds."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402
import torch  # noqa: E402


def us(fn, n=500, warm=30):
    for _ in range(warm):
        fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return round(1e6 * (time.perf_counter() - t) / n, 1)


casc = nv.Cascade(os.path.join(ROOT, "nubomedia-vca_b200", "cascades", "haarcascade_frontalface_alt.xml"))
ctx = nv.Context(0, 1920, 1080)
f1 = torch.from_numpy(synth.frame(640, 480, 4, 1)).pin_memory()
d1 = f1.cuda()
torch.cuda.synchronize()
out = {"host_call": us(lambda: ctx.face_detect(casc, f1.numpy(), 160, 1.25, 3, None))}


def dev():
    ctx.face_submit_device(casc, d1.data_ptr(), 640, 480, 640 * 3, 160, 1.25, 3, None)
    return ctx.face_collect()


out["device_call"] = us(dev)
s = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
x = torch.zeros(1, device="cuda")
with torch.cuda.graph(g, stream=s):
    x += 1


def empty():
    g.replay(); torch.cuda.synchronize()


out["empty_graph_and_sync"] = us(empty)
h = torch.empty(640 * 480 * 3, dtype=torch.uint8).pin_memory()


def h2d():
    d1.view(-1).copy_(h, non_blocking=True); torch.cuda.synchronize()


out["h2d_900KB_and_sync"] = us(h2d)
ctx.set_profile(True)
acc = {}
for i in range(60):
    ctx.face_detect(casc, f1.numpy(), 160, 1.25, 3, None)
    if i >= 10:
        for k, v in ctx.stage_times().items():
            acc.setdefault(k, []).append(v * 1e3)
out["stages_plain_launches"] = {k: round(float(np.median(v)), 1) for k, v in acc.items()}
out["stages_sum"] = round(sum(out["stages_plain_launches"].values()), 1)
print(out)
