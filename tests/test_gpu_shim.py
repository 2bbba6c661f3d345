"""GPU: the reference compiled with the shim (tools/build_shim.py; binaries prebuilt under
shim/_build, they travel with the repository) running on libathena_b200.so -- the reference's
own main(), C++ problem generators, task scheduler and output writers driving the CUDA kernels
through the C ABI -- must reproduce the goldens of the UNMODIFIED reference bit for bit (dt
sequence, restart dumps with ghost zones, history rows).  See tests/shim_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("threads", [1, 2])
def test_reference_with_shim_on_the_gpu(threads):
    import test_shim_cpu
    env = dict(os.environ)
    env.pop("AB_SHIM_LIBDIR", None)
    r = subprocess.run([sys.executable, os.path.join(HERE, "shim_check.py"), "--threads",
                        str(threads)] + test_shim_cpu.SHIM_GOLDENS, env=env, capture_output=True,
                       text=True, timeout=1500)
    print(r.stdout[-4000:], r.stderr[-3000:])
    assert r.returncode == 0 and "shim done: 0 failed" in r.stdout
