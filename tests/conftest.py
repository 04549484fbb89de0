import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    try:  # the torch-based checker must be exact fp32 when it happens to run on the GPU
        import torch

        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:  # noqa: BLE001
        pass


@pytest.fixture(scope="session")
def lib_built():
    """build the C-ABI library (nvcc cross-compiles without a GPU)"""
    import sap3d_build

    return sap3d_build.build()
