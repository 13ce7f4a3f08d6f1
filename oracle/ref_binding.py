"""TEST INFRASTRUCTURE ONLY -- ctypes binding to ``oracle/_ref/libspecbleach_ref.so``.

``libspecbleach_ref.so`` is the UNMODIFIED reference (``/root/reference/src``)
compiled by ``oracle/Makefile`` against the FFTW shim.  It is the strongest
oracle we have: the CUDA path and the numpy restatement
(``oracle/specbleach_np.py``) are both checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

The public entry points bound here are the reference's own
(``include/specbleach_denoiser.h:86-141``, ``include/specbleach_2d_denoiser.h:107-171``);
a few internal module entry points (``nlm_filter_*``) are bound as well so that
single kernels can be checked in isolation.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REF_SO = _HERE / "_ref" / "libspecbleach_ref.so"


class DenoiserParameters(C.Structure):
    """Layout of SpectralBleachDenoiserParameters / SpectralBleach2DDenoiserParameters
    (specbleach_denoiser.h:33-79, specbleach_2d_denoiser.h:37-97): identical field
    order and types in both structs (field 8 is masking_depth in 1D and
    nlm_masking_protection in 2D)."""

    _fields_ = [
        ("learn_noise", C.c_int),
        ("residual_listen", C.c_bool),
        ("reduction_amount", C.c_float),
        ("smoothing_factor", C.c_float),
        ("whitening_factor", C.c_float),
        ("adaptive_noise", C.c_int),
        ("noise_estimation_method", C.c_int),
        ("masking_depth", C.c_float),
        ("suppression_strength", C.c_float),
        ("aggressiveness", C.c_float),
        ("tonal_reduction", C.c_float),
    ]


def make_params(**kw) -> DenoiserParameters:
    p = DenoiserParameters()
    alias = {"nlm_masking_protection": "masking_depth"}
    for k, v in kw.items():
        setattr(p, alias.get(k, k), v)
    return p


class NlmFilterConfig(C.Structure):
    """nlm_filter.h NlmFilterConfig (field order as declared there)."""

    _fields_ = [
        ("spectrum_size", C.c_uint32),
        ("time_buffer_size", C.c_uint32),
        ("patch_size", C.c_uint32),
        ("paste_block_size", C.c_uint32),
        ("search_range_freq", C.c_uint32),
        ("search_range_time_past", C.c_uint32),
        ("search_range_time_future", C.c_uint32),
        ("h_parameter", C.c_float),
        ("distance_threshold", C.c_float),
    ]


def available() -> bool:
    return REF_SO.exists()


_lib = None
_variants: dict = {}

# build variants of the reference (oracle/Makefile, target `variants`): the
# compile-time selectors of the processors pinned to the branches the shipped build never takes
VARIANTS = ("gates", "gss", "sup_global", "sup_bands", "sup_masking", "sup_none", "transient")


def variant_path(name: str) -> Path:
    return _HERE / "_ref" / "variants" / f"libspecbleach_ref_{name}.so"


def load(variant: str | None = None):
    """dlopen the reference build (or one of its build variants) and declare prototypes
    (idempotent per library)."""
    global _lib
    if variant is None and _lib is not None:
        return _lib
    if variant is not None and variant in _variants:
        return _variants[variant]
    path = REF_SO if variant is None else variant_path(variant)
    if not path.exists():
        raise FileNotFoundError(
            f"{path} missing: run `make -C oracle` where /root/reference exists")
    lib = C.CDLL(str(path))
    fp = C.POINTER(C.c_float)
    for pre in ("specbleach_", "specbleach_2d_"):
        g = lambda n: getattr(lib, pre + n)
        g("initialize").restype = C.c_void_p
        g("initialize").argtypes = [C.c_uint32, C.c_float]
        g("free").restype = None
        g("free").argtypes = [C.c_void_p]
        g("load_parameters").restype = C.c_bool
        g("load_parameters").argtypes = [C.c_void_p, DenoiserParameters]
        g("process").restype = C.c_bool
        g("process").argtypes = [C.c_void_p, C.c_uint32, fp, fp]
        g("get_latency").restype = C.c_uint32
        g("get_latency").argtypes = [C.c_void_p]
        g("get_noise_profile_size").restype = C.c_uint32
        g("get_noise_profile_size").argtypes = [C.c_void_p]
        g("load_noise_profile_for_mode").restype = C.c_bool
        g("load_noise_profile_for_mode").argtypes = [
            C.c_void_p, fp, C.c_uint32, C.c_uint32, C.c_int]
        g("reset_noise_profile").restype = C.c_bool
        g("reset_noise_profile").argtypes = [C.c_void_p]
        g("get_noise_profile_block_count_for_mode").restype = C.c_uint32
        g("get_noise_profile_block_count_for_mode").argtypes = [C.c_void_p, C.c_int]
        g("get_noise_profile_for_mode").restype = fp
        g("get_noise_profile_for_mode").argtypes = [C.c_void_p, C.c_int]
        g("noise_profile_available_for_mode").restype = C.c_bool
        g("noise_profile_available_for_mode").argtypes = [C.c_void_p, C.c_int]
    # internal NLM module (nlm_filter.h)
    lib.nlm_filter_initialize.restype = C.c_void_p
    lib.nlm_filter_initialize.argtypes = [NlmFilterConfig]
    lib.nlm_filter_free.restype = None
    lib.nlm_filter_free.argtypes = [C.c_void_p]
    lib.nlm_filter_set_h_parameter.restype = None
    lib.nlm_filter_set_h_parameter.argtypes = [C.c_void_p, C.c_float]
    lib.nlm_filter_push_frame.restype = None
    lib.nlm_filter_push_frame.argtypes = [C.c_void_p, fp]
    lib.nlm_filter_process.restype = C.c_bool
    lib.nlm_filter_process.argtypes = [C.c_void_p, fp]
    if variant is None:
        _lib = lib
    else:
        _variants[variant] = lib
    return lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefDenoiser:
    """Thin OO wrapper over the reference's handle API (1D or 2D)."""

    def __init__(self, sample_rate: int, frame_ms: float, two_d: bool = True, variant: str | None = None):
        self.lib = load(variant)
        self.pre = "specbleach_2d_" if two_d else "specbleach_"
        self.h = self._f("initialize")(int(sample_rate), float(frame_ms))
        if not self.h:
            raise RuntimeError("reference initialize returned NULL")

    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def close(self):
        if self.h:
            self._f("free")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_parameters(self, **kw) -> bool:
        return bool(self._f("load_parameters")(self.h, make_params(**kw)))

    def process(self, x: np.ndarray, chunk: int | None = None) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        n = x.shape[0]
        chunk = n if not chunk else int(chunk)
        pos = 0
        while pos < n:
            m = min(chunk, n - pos)
            ok = self._f("process")(self.h, m, _fp(x[pos:pos + m]), _fp(y[pos:pos + m]))
            assert ok
            pos += m
        return y

    def latency(self) -> int:
        return int(self._f("get_latency")(self.h))

    def profile_size(self) -> int:
        return int(self._f("get_noise_profile_size")(self.h))

    def get_profile(self, mode: int) -> np.ndarray:
        p = self._f("get_noise_profile_for_mode")(self.h, mode)
        n = self.profile_size()
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def load_profile(self, prof: np.ndarray, blocks: int, mode: int) -> bool:
        prof = np.ascontiguousarray(prof, dtype=np.float32)
        return bool(self._f("load_noise_profile_for_mode")(
            self.h, _fp(prof), prof.shape[0], blocks, mode))

    def block_count(self, mode: int) -> int:
        return int(self._f("get_noise_profile_block_count_for_mode")(self.h, mode))

    def profile_available(self, mode: int) -> bool:
        return bool(self._f("noise_profile_available_for_mode")(self.h, mode))

    def reset_profile(self) -> bool:
        return bool(self._f("reset_noise_profile")(self.h))


def ref_nlm(snr: np.ndarray, h: float = 1.0) -> np.ndarray:
    """Run the reference NLM module over a whole SNR map ``snr[F][K]``.

    Frames are pushed one by one (nlm_filter.c:222-236); output row ``j`` is what
    ``nlm_filter_process`` returns after pushing frame ``j`` (target = frame
    ``j-4``), or NaN while the 21-frame ring is not full (nlm_filter.c:238-255).
    """
    lib = load()
    snr = np.ascontiguousarray(snr, dtype=np.float32)
    F, K = snr.shape
    cfg = NlmFilterConfig(K, 21, 8, 8, 8, 16, 4, 1.0, 0.0)
    f = lib.nlm_filter_initialize(cfg)
    lib.nlm_filter_set_h_parameter(f, float(h))
    out = np.full((F, K), np.nan, dtype=np.float32)
    row = np.empty(K, dtype=np.float32)
    for j in range(F):
        lib.nlm_filter_push_frame(f, _fp(snr[j]))
        if lib.nlm_filter_process(f, _fp(row)):
            out[j] = row
    lib.nlm_filter_free(f)
    return out


def set_threads(n: int) -> None:
    """The reference reads OMP_NUM_THREADS at nlm_filter_initialize time
    (nlm_filter.c:127-136)."""
    os.environ["OMP_NUM_THREADS"] = str(int(n))
