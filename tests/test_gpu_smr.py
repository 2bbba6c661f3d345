"""GPU: static mesh refinement on the device path (hydro + passive scalars, one process), in its
own process so that a CUDA error there cannot poison the context of the other tests."""
import os
import subprocess
import sys

import pytest

import util

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def device_smr_goldens():
    """the SMR fixtures inside the device path's scope: MeshBlocks of at least 2*NGHOST cells"""
    out = []
    for n in util.golden_names(include_smr=True):
        if not n.startswith("smr_"):
            continue
        g = util.Golden(n)
        mb = g.par.get("meshblock", {})
        sizes = [int(mb.get("nx%d" % d, 1)) for d in (1, 2, 3)]
        if all(s == 1 or s >= 2*g.ng for s in sizes):
            out.append(n)
    return out


def test_refined_meshes_reproduce_reference_goldens():
    names = device_smr_goldens()
    assert len(names) >= 5
    r = subprocess.run([sys.executable, os.path.join(HERE, "smr_check.py")] + names,
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-3000:])
    assert r.returncode == 0 and "smr done: 0 failed" in r.stdout
