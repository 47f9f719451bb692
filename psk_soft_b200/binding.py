"""ctypes binding of include/pskd.h -- the only way Python reaches the CUDA path.

Loading fails loudly (ImportError / RuntimeError) when libpskd.so is missing or stale and cannot
be rebuilt; there is no Python or CPU implementation behind this module.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

PSKD_OK = 0
PSKD_IGNORED_REAL_DATA = 1
PSKD_NO_BITS = 2
PSKD_ERR_ARG, PSKD_ERR_CUDA, PSKD_ERR_NOMEM, PSKD_ERR_UNSUPPORTED, PSKD_ERR_CAPACITY = -1, -2, -3, -4, -5
FLAG_HOST_BUFFERS, FLAG_QUEUE_FLUSHED, FLAG_SRI_CHANGED, FLAG_NO_SYNC = 1, 2, 4, 8

# every symbol include/pskd.h declares (tests check the library exports all of them)
EXPORTS = ("pskd_default_props", "pskd_create", "pskd_destroy", "pskd_set_props", "pskd_get_props",
           "pskd_max_symbols", "pskd_process", "pskd_sync", "pskd_stream", "pskd_get_sri", "pskd_get_stats",
           "pskd_launch_count", "pskd_last_error", "pskd_abi_version", "pskd_device_count", "pskd_synth_fill",
           "pskd_profile_enable", "pskd_profile_read", "pskd_state_size", "pskd_state_export", "pskd_state_import")


class Props(C.Structure):
    _fields_ = [("samplesPerBaud", C.c_uint16), ("numAvg", C.c_uint32), ("constelationSize", C.c_uint16),
                ("phaseAvg", C.c_uint16), ("differentialDecoding", C.c_uint8), ("resetState", C.c_uint8)]


class Input(C.Structure):
    _fields_ = [("iq", C.c_void_p), ("iq_stride", C.c_size_t), ("n_complex", C.POINTER(C.c_size_t)),
                ("n_complex_all", C.c_size_t), ("sri_xdelta", C.c_double), ("sri_mode", C.c_int),
                ("packet_len", C.c_size_t), ("flags", C.c_uint32)]


class Output(C.Structure):
    _fields_ = [("soft", C.c_void_p), ("bits", C.c_void_p), ("phase", C.c_void_p), ("sample_index", C.c_void_p),
                ("sym_stride", C.c_size_t), ("bits_stride", C.c_size_t),
                ("n_symbols", C.POINTER(C.c_size_t)), ("n_bits", C.POINTER(C.c_size_t)), ("hard", C.c_void_p)]


class SriOut(C.Structure):
    _fields_ = [("soft_xdelta", C.c_double), ("soft_mode", C.c_int), ("phase_xdelta", C.c_double), ("phase_mode", C.c_int),
                ("bits_xdelta", C.c_double), ("bits_mode", C.c_int), ("sri_pushes", C.c_long)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("symbols_out", "samples_in", "packets", "wraps", "spec_chunks",
                                          "spec_misses", "seq_channels", "tp_packets", "tp_repaired")]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ms_total", C.c_double), ("launches", C.c_uint64), ("alg_bytes", C.c_double)]


class Synth(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("samplesPerBaud", C.c_uint16), ("constelationSize", C.c_uint16),
                ("sigma", C.c_float), ("freq_max", C.c_float), ("pn_sigma", C.c_float), ("shaped", C.c_float),
                ("period", C.c_uint32)]


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """dlopen libpskd.so (building it first when sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.isfile(_build.LIB_PATH):
        raise ImportError(f"{_build.LIB_PATH} is missing: the CUDA extension was not built "
                          "(run `python -m psk_soft_b200._build`); there is no CPU fallback")
    lib = C.CDLL(os.environ.get("PSKD_LIB") or _build.LIB_PATH)   # PSKD_LIB: tuning builds only
    H = C.c_void_p
    lib.pskd_default_props.argtypes = [C.POINTER(Props)]; lib.pskd_default_props.restype = None
    lib.pskd_create.argtypes = [C.POINTER(H), C.c_int, C.c_int, C.POINTER(Props)]; lib.pskd_create.restype = C.c_int
    lib.pskd_destroy.argtypes = [H]; lib.pskd_destroy.restype = C.c_int
    lib.pskd_set_props.argtypes = [H, C.c_int, C.POINTER(Props)]; lib.pskd_set_props.restype = C.c_int
    lib.pskd_get_props.argtypes = [H, C.c_int, C.POINTER(Props)]; lib.pskd_get_props.restype = C.c_int
    lib.pskd_max_symbols.argtypes = [H, C.c_int, C.c_size_t]; lib.pskd_max_symbols.restype = C.c_size_t
    lib.pskd_process.argtypes = [H, C.POINTER(Input), C.POINTER(Output)]; lib.pskd_process.restype = C.c_int
    lib.pskd_sync.argtypes = [H]; lib.pskd_sync.restype = C.c_int
    lib.pskd_stream.argtypes = [H]; lib.pskd_stream.restype = C.c_void_p
    lib.pskd_get_sri.argtypes = [H, C.c_int, C.POINTER(SriOut)]; lib.pskd_get_sri.restype = C.c_int
    lib.pskd_get_stats.argtypes = [H, C.POINTER(Stats)]; lib.pskd_get_stats.restype = C.c_int
    lib.pskd_launch_count.argtypes = [H]; lib.pskd_launch_count.restype = C.c_uint64
    lib.pskd_last_error.argtypes = []; lib.pskd_last_error.restype = C.c_char_p
    lib.pskd_abi_version.argtypes = []; lib.pskd_abi_version.restype = C.c_int
    lib.pskd_device_count.argtypes = []; lib.pskd_device_count.restype = C.c_int
    lib.pskd_profile_enable.argtypes = [H, C.c_int]; lib.pskd_profile_enable.restype = C.c_int
    lib.pskd_profile_read.argtypes = [H, C.POINTER(KernelTime), C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.pskd_profile_read.restype = C.c_int
    lib.pskd_state_size.argtypes = [H]; lib.pskd_state_size.restype = C.c_size_t
    lib.pskd_state_export.argtypes = [H, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]; lib.pskd_state_export.restype = C.c_int
    lib.pskd_state_import.argtypes = [H, C.c_void_p, C.c_size_t]; lib.pskd_state_import.restype = C.c_int
    lib.pskd_synth_fill.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.POINTER(Synth), C.c_void_p]
    lib.pskd_synth_fill.restype = C.c_int
    _lib = lib
    return lib


class PskdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pskd error {code}: {msg}")
        self.code = code
