import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must FAIL (not skip) on a box without the GPU / the CUDA library:
    # the product has no CPU fallback.  Nothing to do here; the marker is only a selector.
    pass


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library (nvcc cross-compiles without a GPU) and the C oracle once per session when
    they are missing or stale; the built files are git-ignored."""
    import shutil
    from gnomix_b200 import build as b
    if b.is_stale() and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        b.build_lib()
    from oracle import c_oracle
    c_oracle.build()


@pytest.fixture(scope="session")
def libgnx(_built):
    from gnomix_b200 import _lib
    return _lib.lib()
