import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_HAVE_GPU = None


def _gpu_available():
    """One probe of ncme_ctx_create per session (no torch import, no CPU fallback to hide behind)."""
    global _HAVE_GPU
    if _HAVE_GPU is None:
        try:
            import __graft_entry__ as g
            g.load_package().Context.default()
            _HAVE_GPU = True
        except Exception:
            _HAVE_GPU = False
    return _HAVE_GPU


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only box: skip (not fail) the gpu-marked tests so CPU regressions stay visible."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _gpu_available():
        skip = pytest.mark.skip(reason="no CUDA device available (libncme has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    if not _gpu_available():
        pytest.skip("no CUDA device available")
    return pkg.Context.default()
