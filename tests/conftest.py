import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One rt_ctx for the whole GPU session (fails loudly when the CUDA library or the device is missing)."""
    from ilgpu_raytracing_b200 import build, native
    build.build_core()
    ctx = native.Context(0)
    yield ctx
    ctx.close()
