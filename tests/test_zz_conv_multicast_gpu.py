"""The 2-CTA-cluster persistent convolution with multicast weight tiles (conv_tc_persist_mc_kernel, opt-in through
SAP3D_CONV_MULTICAST=2|4 = cluster size, read once per process): the persistent-kernel parity cases and the full-size linearity property of
tests/test_conv_gpu.py must hold with it switched on.  Runs them in a child process so the switch is seen at first use."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cluster", ["2", "4"])
def test_conv_parity_with_multicast_weight_tiles(lib_built, cluster):
    env = dict(os.environ, SAP3D_CONV_MULTICAST=cluster)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_conv_gpu.py"), "-q", "-x", "-m", "gpu",
                        "-k", "persistent or linearity or tensor_core_conv"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
