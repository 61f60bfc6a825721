"""Three config-2 frames (eyes inside faces on a 1280x720 frame, stand-in eye model) through the nuboeyedetector mirror: the
command for the launch list of the nested-element path."""
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nubomedia-vca_b200", "python"))
import nubovca as nv  # noqa: E402
from nubovca import synth  # noqa: E402

src = os.path.join(ROOT, "nubomedia-vca_b200", "cascades")
d = tempfile.mkdtemp()
shutil.copy(os.path.join(src, "haarcascade_frontalface_alt.xml"), d)
for f in ("haarcascade_mcs_lefteye.xml", "haarcascade_mcs_righteye.xml"):
    shutil.copy(os.path.join(src, "haarcascade_eye.xml"), os.path.join(d, f))
f2 = synth.frame(1280, 720, 3, 2, smin=0.4, smax=0.6)
e = nv.Element("nuboeyedetector", 0, d)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    r = e.process(f2)
print(r)
e.close()
shutil.rmtree(d, ignore_errors=True)
