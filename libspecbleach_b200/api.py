"""ctypes mirror of the C ABI in ``include/specbleach_denoiser.h``,
``include/specbleach_2d_denoiser.h`` and ``include/specbleach_b200.h``.

Function names, argument order and error behaviour are the reference's
(``src/processors/specbleach_denoiser.c``, ``specbleach_2d_denoiser.c``): NULL /
False / 0 on error, parameters struct passed by value.  ``Denoiser`` is a thin
object wrapper so that parity tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libspecbleach_b200.so"


class DenoiserParameters(C.Structure):
    """SpectralBleachDenoiserParameters == SpectralBleach2DDenoiserParameters layout."""

    _fields_ = [
        ("learn_noise", C.c_int),
        ("residual_listen", C.c_bool),
        ("reduction_amount", C.c_float),
        ("smoothing_factor", C.c_float),
        ("whitening_factor", C.c_float),
        ("adaptive_noise", C.c_int),
        ("noise_estimation_method", C.c_int),
        ("masking_depth", C.c_float),          # nlm_masking_protection in the 2D struct
        ("suppression_strength", C.c_float),
        ("aggressiveness", C.c_float),
        ("tonal_reduction", C.c_float),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64),
        ("nlm_launches", C.c_uint64),
        ("nlm_ms", C.c_double),
        ("nlm_frame_blocks", C.c_uint64),
        ("frames", C.c_uint64),
    ]


def make_params(**kw) -> DenoiserParameters:
    p = DenoiserParameters()
    alias = {"nlm_masking_protection": "masking_depth"}
    for k, v in kw.items():
        setattr(p, alias.get(k, k), v)
    return p


_lib = None
_fp = C.POINTER(C.c_float)

PUBLIC_SYMBOLS = [pre + n for pre in ("specbleach_", "specbleach_2d_") for n in (
    "initialize", "free", "load_parameters", "process", "get_latency",
    "get_noise_profile_size", "load_noise_profile_for_mode", "reset_noise_profile",
    "get_noise_profile_block_count_for_mode", "get_noise_profile_for_mode",
    "noise_profile_available_for_mode")]
EXTENSION_SYMBOLS = ["specbleach_b200_" + n for n in (
    "device_count", "set_device", "last_error", "host_alloc", "host_free", "configure",
    "process_device", "process_batch", "process_batch_device", "profile", "stats_reset",
    "stats_get", "stats_classes", "stats_class_name", "stats_class", "get_geometry",
    "debug_fp32_peak_tflops", "debug_check_guards", "debug_nlm", "debug_rfft", "profile_save", "profile_load",
    "state_size", "state_save", "state_load", "process_interleaved", "process_tiled", "process_tiled_batch", "set_strategy")]


def lib():
    """dlopen ``libspecbleach_b200.so`` and declare prototypes (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not built: run `make -C libspecbleach_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    for pre in ("specbleach_", "specbleach_2d_"):
        g = lambda n: getattr(L, pre + n)
        g("initialize").restype = C.c_void_p
        g("initialize").argtypes = [C.c_uint32, C.c_float]
        g("free").restype = None
        g("free").argtypes = [C.c_void_p]
        g("load_parameters").restype = C.c_bool
        g("load_parameters").argtypes = [C.c_void_p, DenoiserParameters]
        g("process").restype = C.c_bool
        g("process").argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        g("get_latency").restype = C.c_uint32
        g("get_latency").argtypes = [C.c_void_p]
        g("get_noise_profile_size").restype = C.c_uint32
        g("get_noise_profile_size").argtypes = [C.c_void_p]
        g("load_noise_profile_for_mode").restype = C.c_bool
        g("load_noise_profile_for_mode").argtypes = [
            C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        g("reset_noise_profile").restype = C.c_bool
        g("reset_noise_profile").argtypes = [C.c_void_p]
        g("get_noise_profile_block_count_for_mode").restype = C.c_uint32
        g("get_noise_profile_block_count_for_mode").argtypes = [C.c_void_p, C.c_int]
        g("get_noise_profile_for_mode").restype = _fp
        g("get_noise_profile_for_mode").argtypes = [C.c_void_p, C.c_int]
        g("noise_profile_available_for_mode").restype = C.c_bool
        g("noise_profile_available_for_mode").argtypes = [C.c_void_p, C.c_int]
    L.specbleach_b200_device_count.restype = C.c_int
    L.specbleach_b200_set_device.restype = C.c_bool
    L.specbleach_b200_set_device.argtypes = [C.c_int]
    L.specbleach_b200_last_error.restype = C.c_char_p
    L.specbleach_b200_host_alloc.restype = C.c_void_p
    L.specbleach_b200_host_alloc.argtypes = [C.c_size_t]
    L.specbleach_b200_host_free.restype = None
    L.specbleach_b200_host_free.argtypes = [C.c_void_p]
    L.specbleach_b200_configure.restype = C.c_bool
    L.specbleach_b200_configure.argtypes = [C.c_uint32, C.c_uint32]
    L.specbleach_b200_process_device.restype = C.c_bool
    L.specbleach_b200_process_device.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    for n in ("process_batch", "process_batch_device"):
        f = getattr(L, "specbleach_b200_" + n)
        f.restype = C.c_bool
        f.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(C.c_uint32),
                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.specbleach_b200_profile.restype = None
    L.specbleach_b200_profile.argtypes = [C.c_int]
    L.specbleach_b200_stats_reset.restype = None
    L.specbleach_b200_stats_get.restype = C.c_bool
    L.specbleach_b200_stats_get.argtypes = [C.POINTER(Stats)]
    L.specbleach_b200_stats_classes.restype = C.c_int
    L.specbleach_b200_stats_class_name.restype = C.c_char_p
    L.specbleach_b200_stats_class_name.argtypes = [C.c_int]
    L.specbleach_b200_stats_class.restype = C.c_bool
    L.specbleach_b200_stats_class.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.specbleach_b200_profile_save.restype = C.c_bool
    L.specbleach_b200_profile_save.argtypes = [C.c_void_p, C.c_char_p]
    L.specbleach_b200_profile_load.restype = C.c_bool
    L.specbleach_b200_profile_load.argtypes = [C.c_void_p, C.c_char_p]
    L.specbleach_b200_state_size.restype = C.c_size_t
    L.specbleach_b200_state_size.argtypes = [C.c_void_p]
    L.specbleach_b200_state_save.restype = C.c_bool
    L.specbleach_b200_state_save.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.specbleach_b200_state_load.restype = C.c_bool
    L.specbleach_b200_state_load.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.specbleach_b200_process_interleaved.restype = C.c_bool
    L.specbleach_b200_process_interleaved.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32,
                                                      C.c_int, C.c_void_p, C.c_void_p]
    L.specbleach_b200_process_tiled.restype = C.c_bool
    L.specbleach_b200_process_tiled.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                                C.c_uint32, C.POINTER(C.c_int), C.c_uint32, C.c_uint32]
    L.specbleach_b200_process_tiled_batch.restype = C.c_bool
    L.specbleach_b200_process_tiled_batch.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(C.c_uint32),
                                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint32,
                                                      C.c_uint32, C.POINTER(C.c_int), C.c_uint32, C.c_uint32]
    L.specbleach_b200_set_strategy.restype = C.c_bool
    L.specbleach_b200_set_strategy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.specbleach_b200_get_geometry.restype = C.c_bool
    L.specbleach_b200_get_geometry.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 4
    L.specbleach_b200_debug_fp32_peak_tflops.restype = C.c_double
    L.specbleach_b200_debug_check_guards.restype = C.c_int
    L.specbleach_b200_debug_nlm.restype = C.c_bool
    L.specbleach_b200_debug_nlm.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float,
                                            C.c_void_p]
    L.specbleach_b200_debug_rfft.restype = C.c_bool
    L.specbleach_b200_debug_rfft.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
    _lib = L
    return L


def last_error() -> str:
    return lib().specbleach_b200_last_error().decode()


def device_count() -> int:
    return int(lib().specbleach_b200_device_count())


def set_device(idx: int) -> bool:
    return bool(lib().specbleach_b200_set_device(int(idx)))


def stats() -> dict:
    s = Stats()
    if not lib().specbleach_b200_stats_get(C.byref(s)):
        raise RuntimeError(last_error())
    return {k: getattr(s, k) for k, _ in Stats._fields_}


def class_stats() -> dict:
    """per-kernel-class {name: (ms, launches)} measured with CUDA events (profile mode)."""
    L = lib()
    out = {}
    for c in range(L.specbleach_b200_stats_classes()):
        ms, n = C.c_double(), C.c_uint64()
        L.specbleach_b200_stats_class(c, C.byref(ms), C.byref(n))
        if n.value:
            out[L.specbleach_b200_stats_class_name(c).decode()] = (ms.value, n.value)
    return out


class PinnedArray:
    """float32 numpy view over a page-locked buffer from specbleach_b200_host_alloc."""

    def __init__(self, n: int):
        self.n = int(n)
        self.ptr = lib().specbleach_b200_host_alloc(self.n * 4)
        if not self.ptr:
            raise MemoryError(last_error())
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, _fp), shape=(self.n,))

    def free(self):
        if self.ptr:
            lib().specbleach_b200_host_free(self.ptr)
            self.ptr = None
            self.array = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Denoiser:
    """Object wrapper over one handle of the drop-in API (1D or 2D)."""

    def __init__(self, sample_rate: int, frame_ms: float, two_d: bool = True):
        self.L = lib()
        self.pre = "specbleach_2d_" if two_d else "specbleach_"
        self.two_d = two_d
        self.h = self._f("initialize")(int(sample_rate), float(frame_ms))
        if not self.h:
            raise RuntimeError(f"{self.pre}initialize returned NULL: {last_error()}")

    def _f(self, name):
        return getattr(self.L, self.pre + name)

    def close(self):
        if getattr(self, "h", None):
            self._f("free")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_parameters(self, **kw) -> bool:
        return bool(self._f("load_parameters")(self.h, make_params(**kw)))

    def process_raw(self, n: int, in_ptr, out_ptr) -> bool:
        return bool(self._f("process")(self.h, int(n), in_ptr, out_ptr))

    def process(self, x: np.ndarray, chunk: int | None = None) -> np.ndarray:
        """specbleach[_2d]_process over ``x`` in calls of ``chunk`` samples (host buffers)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        n = x.shape[0]
        chunk = n if not chunk else int(chunk)
        pos = 0
        while pos < n:
            m = min(chunk, n - pos)
            if not self.process_raw(m, x[pos:].ctypes.data, y[pos:].ctypes.data):
                raise RuntimeError("process failed: " + last_error())
            pos += m
        return y

    def latency(self) -> int:
        return int(self._f("get_latency")(self.h))

    def profile_size(self) -> int:
        return int(self._f("get_noise_profile_size")(self.h))

    def get_profile(self, mode: int) -> np.ndarray:
        p = self._f("get_noise_profile_for_mode")(self.h, mode)
        return np.ctypeslib.as_array(p, shape=(self.profile_size(),)).copy()

    def load_profile(self, prof: np.ndarray, blocks: int, mode: int) -> bool:
        prof = np.ascontiguousarray(prof, dtype=np.float32)
        return bool(self._f("load_noise_profile_for_mode")(
            self.h, prof.ctypes.data, prof.shape[0], blocks, mode))

    def block_count(self, mode: int) -> int:
        return int(self._f("get_noise_profile_block_count_for_mode")(self.h, mode))

    def profile_available(self, mode: int) -> bool:
        return bool(self._f("noise_profile_available_for_mode")(self.h, mode))

    def reset_profile(self) -> bool:
        return bool(self._f("reset_noise_profile")(self.h))

    # ---- extensions: persistence (include/specbleach_b200.h) ----
    def save_profile(self, path) -> bool:
        return bool(self.L.specbleach_b200_profile_save(self.h, str(path).encode()))

    def restore_profile(self, path) -> bool:
        return bool(self.L.specbleach_b200_profile_load(self.h, str(path).encode()))

    def save_state(self) -> bytes:
        n = int(self.L.specbleach_b200_state_size(self.h))
        buf = C.create_string_buffer(n)
        if not self.L.specbleach_b200_state_save(self.h, buf, n):
            raise RuntimeError("state_save failed: " + last_error())
        return buf.raw

    def load_state(self, blob: bytes) -> bool:
        return bool(self.L.specbleach_b200_state_load(self.h, blob, len(blob)))

    def process_tiled(self, x: np.ndarray, tiles: int, warmup_frames: int = 512, devices=None,
                      allow_adaptive: bool = False, out: np.ndarray | None = None) -> np.ndarray:
        """specbleach_b200_process_tiled: one call over ``x`` cut into ``tiles`` time tiles."""
        x = np.ascontiguousarray(x, dtype=np.float32) if not isinstance(x, np.ndarray) or x.dtype != np.float32 else x
        y = np.empty_like(x) if out is None else out
        devs = (C.c_int * len(devices))(*devices) if devices else None
        if not self.L.specbleach_b200_process_tiled(self.h, x.shape[0], x.ctypes.data, y.ctypes.data, int(tiles),
                                                    int(warmup_frames), devs, len(devices) if devices else 0,
                                                    1 if allow_adaptive else 0):
            raise RuntimeError("process_tiled failed: " + last_error())
        return y

    def set_strategy(self, gain_type: int = 0, suppression_type: int = 0, smoothing_type: int = 1) -> bool:
        """specbleach_b200_set_strategy: the reference's compile-time strategy selectors (N4)."""
        return bool(self.L.specbleach_b200_set_strategy(self.h, gain_type, suppression_type, smoothing_type))

    def geometry(self) -> dict:
        v = [C.c_uint32() for _ in range(4)]
        self.L.specbleach_b200_get_geometry(self.h, *[C.byref(a) for a in v])
        return dict(zip(("frame", "fft_size", "hop", "bins"), (int(a.value) for a in v)))


class Denoiser2D(Denoiser):
    def __init__(self, sample_rate: int, frame_ms: float):
        super().__init__(sample_rate, frame_ms, True)


class Denoiser1D(Denoiser):
    def __init__(self, sample_rate: int, frame_ms: float):
        super().__init__(sample_rate, frame_ms, False)


def process_batch(denoisers, inputs, outputs, device: bool = False) -> bool:
    """specbleach_b200_process_batch[_device] over parallel lists.

    ``inputs``/``outputs`` are numpy float32 arrays (host) or integer device
    pointers with ``(ptr, n)`` tuples when ``device`` is True."""
    L = lib()
    cnt = len(denoisers)
    hs = (C.c_void_p * cnt)(*[d.h for d in denoisers])
    if device:
        ns = (C.c_uint32 * cnt)(*[int(n) for _, n in inputs])
        ins = (C.c_void_p * cnt)(*[int(p) for p, _ in inputs])
        outs = (C.c_void_p * cnt)(*[int(p) for p, _ in outputs])
        ok = L.specbleach_b200_process_batch_device(hs, cnt, ns, ins, outs)
    else:
        ns = (C.c_uint32 * cnt)(*[int(a.shape[0]) for a in inputs])
        ins = (C.c_void_p * cnt)(*[a.ctypes.data for a in inputs])
        outs = (C.c_void_p * cnt)(*[a.ctypes.data for a in outputs])
        ok = L.specbleach_b200_process_batch(hs, cnt, ns, ins, outs)
    return bool(ok)


def process_tiled_batch(denoisers, inputs, outputs, tiles: int, warmup_frames: int = 512, devices=None,
                        allow_adaptive: bool = False) -> bool:
    """specbleach_b200_process_tiled_batch over parallel lists of host float32 arrays."""
    L = lib()
    cnt = len(denoisers)
    hs = (C.c_void_p * cnt)(*[d.h for d in denoisers])
    ns = (C.c_uint32 * cnt)(*[int(a.shape[0]) for a in inputs])
    ins = (C.c_void_p * cnt)(*[a.ctypes.data for a in inputs])
    outs = (C.c_void_p * cnt)(*[a.ctypes.data for a in outputs])
    devs = (C.c_int * len(devices))(*devices) if devices else None
    return bool(L.specbleach_b200_process_tiled_batch(hs, cnt, ns, ins, outs, int(tiles), int(warmup_frames), devs,
                                                      len(devices) if devices else 0, 1 if allow_adaptive else 0))


PCM_FORMATS = {"f32": 0, "s16": 1, "s24": 2, "s32": 3}


def process_interleaved(denoisers, pcm: np.ndarray, fmt: str) -> np.ndarray:
    """specbleach_b200_process_interleaved: ``pcm`` is [frames, channels] (float32, int16,
    int32) or, for "s24", a uint8 array of frames*channels*3 packed bytes."""
    L = lib()
    ch = len(denoisers)
    pcm = np.ascontiguousarray(pcm)
    frames = pcm.size // (3 * ch) if fmt == "s24" else pcm.shape[0]
    out = np.empty_like(pcm)
    hs = (C.c_void_p * ch)(*[d.h for d in denoisers])
    if not L.specbleach_b200_process_interleaved(hs, ch, frames, PCM_FORMATS[fmt], pcm.ctypes.data,
                                                 out.ctypes.data):
        raise RuntimeError("process_interleaved failed: " + last_error())
    return out


def debug_nlm(snr: np.ndarray, h: float = 1.0) -> np.ndarray:
    snr = np.ascontiguousarray(snr, dtype=np.float32)
    out = np.full_like(snr, np.nan)
    if not lib().specbleach_b200_debug_nlm(snr.ctypes.data, snr.shape[0], snr.shape[1],
                                           float(h), out.ctypes.data):
        raise RuntimeError("debug_nlm failed: " + last_error())
    return out


def debug_rfft(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    frames, n = x.shape
    hc = np.empty_like(x)
    rt = np.empty_like(x)
    if not lib().specbleach_b200_debug_rfft(n, frames, x.ctypes.data, hc.ctypes.data,
                                            rt.ctypes.data):
        raise RuntimeError("debug_rfft failed: " + last_error())
    return hc, rt
