import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present() -> bool:
    """True iff the product library loads and finds a CUDA device (og_init succeeds on device 0)."""
    try:
        import owshen_b200 as ob
        c = ob.Context(0)
        c.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a host without a GPU skips the gpu-marked tests instead of erroring in their fixtures
    (the product has no CPU path to fall back to, so there is nothing for them to test there)."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device (libowshen_b200.so has no CPU path)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def ctx():
    import owshen_b200 as ob
    try:
        c = ob.Context(0)
    except (OSError, ob.OwshenB200Error) as e:
        pytest.skip(f"no CUDA device: {e}")
    yield c
    c.close()
