import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rb.load()
    return rb


@pytest.fixture(scope="session")
def gpu():
    import libspecbleach_b200 as sb
    if sb.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    sb.set_device(0)
    return sb


@pytest.fixture(scope="session", autouse=True)
def guard_zones_intact_at_exit():
    """`SB_GUARD=1 python -m pytest tests -m gpu` runs the whole GPU suite with guard zones
    around every device allocation (include/specbleach_b200.h); none may be written."""
    yield
    import os
    if os.environ.get("SB_GUARD", "0") not in ("", "0", "selftest"):
        from libspecbleach_b200 import api
        assert api.lib().specbleach_b200_debug_check_guards() == 0
