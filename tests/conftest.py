import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    # make sure the in-tree artefacts exist (nvcc cross-compiles without a GPU; seconds per file)
    from cadr_b200 import build
    build.build_cuda()
    build.build_oracle()
    build.build_host()


def _has_gpu() -> bool:
    import cadr_b200
    try:
        c = cadr_b200.Context(0)
        c.close()
        return True
    except cadr_b200.CadrError:
        return False


@pytest.fixture(scope="session")
def ctx():
    """A context on cuda:0.  GPU tests FAIL (not skip) when the library cannot reach a device: a silent
    fallback would void every parity claim."""
    import cadr_b200
    c = cadr_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def address_ctx():
    import cadr_b200
    c = cadr_b200.Context(None)
    yield c
    c.close()
