import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    # make sure the in-tree libraries exist (no-op when up to date)
    from icsfoam_b200 import build
    build.build_meshtools()
    build.build_oracle()
    if not os.path.exists(os.path.join(ROOT, "icsfoam_b200", "libicsb200.so")):
        build.build_product()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture
def gpu_context():
    """A CUDA context of the product.  Fails loudly (no skip, no fallback) when the extension cannot run."""
    from icsfoam_b200.context import Context
    ctxs = []

    def make(**kw):
        c = Context(**kw)
        ctxs.append(c)
        return c

    yield make
    for c in ctxs:
        c.close()
