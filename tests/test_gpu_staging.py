"""GPU: pipelined host <-> device staging (ab_stage_*), in its own process so that a CUDA error
there cannot poison the context of the other tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_staged_pipeline_matches_plain_sequence():
    r = subprocess.run([sys.executable, os.path.join(HERE, "stage_check.py")],
                       capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "staging ok" in r.stdout
