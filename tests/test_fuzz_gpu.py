"""A short run of tools/fuzz_parity.py: random sizes, strides, parameters and input formats through long-lived contexts
against the oracle (the full-length runs of the round, 3843 cases, are recorded in DESIGN.md)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_short_randomised_parity_run():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "15", "7"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout
