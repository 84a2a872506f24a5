import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from box2d_optimized_b200 import capi
        return capi.load_cuda().b2g_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: a skipped parity test is a
    # silent fallback.  Without -m gpu (CPU container) the gpu tests are deselected by the marker.
    pass


@pytest.fixture(scope="session")
def ref_available():
    import oracle.bindings as oracle_bindings
    return os.path.exists(oracle_bindings.ref_path())


@pytest.fixture(scope="session")
def require_ref(ref_available):
    if not ref_available:
        pytest.fail("oracle/_ref/libb2ref.so is missing: build it with `make -C oracle ref` where "
                    "/root/reference exists (it travels to the GPU box with the snapshot)")
