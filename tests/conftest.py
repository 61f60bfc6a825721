import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "nubomedia-vca_b200", "python"),
          os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASCADE_DIR = os.path.join(ROOT, "nubomedia-vca_b200", "cascades")

# a fresh checkout has no built artefacts (they are git-ignored): build them once, as __graft_entry__.build() does
_LIB = os.path.join(ROOT, "nubomedia-vca_b200", "lib")
if not (os.path.exists(os.path.join(_LIB, "libnubovca.so")) and os.path.exists(os.path.join(_LIB, "streams_bench"))):
    import __graft_entry__
    __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cascade_dir():
    return CASCADE_DIR
