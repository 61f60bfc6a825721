"""A short nubotracker sequence (BASELINE config 4: 1280x720 BGRA) on one context: the command the ncu captures of the
tracker kernels in profiles/ were taken with (run it without ncu first)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402

ctx = nv.Context(0, 1280, 720)
n = 0
for i, f in enumerate(synth.tracker_sequence(1280, 720, 6, seed=4)):
    n += len(ctx.tracker_process(f, 33.3 * (i + 1)))
print("objects", n, ctx.counters())
ctx.close()
