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
