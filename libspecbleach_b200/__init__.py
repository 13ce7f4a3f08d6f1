"""specbleach-b200: B200-native implementation of libspecbleach's offline denoising
hot path behind the reference's own C API.

The product is ``libspecbleach_b200.so`` (CUDA, sm_100a, C ABI declared in
``include/*.h``).  This package is only the Python-side mirror of that ABI used by
the tests and ``bench.py``; it never computes anything itself and there is no CPU
fallback: importing :mod:`libspecbleach_b200.api` fails loudly when the shared
library has not been built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
from .api import (  # noqa: F401
    Denoiser,
    Denoiser1D,
    Denoiser2D,
    DenoiserParameters,
    LIB_PATH,
    device_count,
    last_error,
    lib,
    make_params,
    process_batch,
    set_device,
)

__all__ = [
    "Denoiser", "Denoiser1D", "Denoiser2D", "DenoiserParameters", "LIB_PATH", "device_count",
    "last_error", "lib", "make_params", "process_batch", "set_device",
]
