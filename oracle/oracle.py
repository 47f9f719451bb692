"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends for the two CPU checkers.

* ``RefComponent``    : the UNMODIFIED reference ``psk_soft_i`` (reference: cpp/psk_soft.cpp)
                        compiled in place into ``oracle/_ref/libpsk_ref.so`` (see oracle/Makefile).
* ``OracleComponent`` : our plain-C restatement ``oracle/psk_oracle.c`` ->
                        ``oracle/_build/libpsk_oracle.so``.

Both expose the same packet-level interface (configure / push / demod), mirroring what the
REDHAWK sandbox does to the component in the reference's own test
(reference: tests/test_psk_soft.py:241-269).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  Nothing under ``psk_soft_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpsk_ref.so")
ORC_SO = os.path.join(HERE, "_build", "libpsk_oracle.so")
REFERENCE_ROOT = "/root/reference"

PROPS = ("samplesPerBaud", "numAvg", "constelationSize", "phaseAvg", "differentialDecoding", "resetState")


def build(ref: bool | None = None) -> None:
    """Compile the C restatement (always) and the reference build (when its sources exist)."""
    targets = ["oracle"]
    have_ref_src = os.path.isfile(os.path.join(REFERENCE_ROOT, "cpp", "psk_soft.cpp"))
    if ref is None:
        ref = have_ref_src
    if ref and have_ref_src:
        targets.append("ref")
    subprocess.check_call(["make", "-s", "-C", HERE] + targets)


def have_ref() -> bool:
    return os.path.isfile(REF_SO)


def _bind(lib, prefix):
    f = lambda n: getattr(lib, prefix + n)
    f("create").restype = C.c_void_p
    f("create").argtypes = []
    f("destroy").argtypes = [C.c_void_p]
    f("destroy").restype = None
    f("configure").argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    f("configure").restype = C.c_int
    f("query").argtypes = [C.c_void_p, C.c_char_p]
    f("query").restype = C.c_double
    f("push").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_int]
    f("push").restype = C.c_int
    f("read").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    f("read").restype = C.c_size_t
    f("sri_count").argtypes = [C.c_void_p, C.c_int]
    f("sri_count").restype = C.c_long
    f("sri_xdelta").argtypes = [C.c_void_p, C.c_int]
    f("sri_xdelta").restype = C.c_double
    f("sri_mode").argtypes = [C.c_void_p, C.c_int]
    f("sri_mode").restype = C.c_int
    f("packet_count").argtypes = [C.c_void_p, C.c_int]
    f("packet_count").restype = C.c_long
    f("demod").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_double,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
    f("demod").restype = C.c_size_t
    return f


class _Component:
    """Packet-level driver shared by the reference build and the C restatement."""
    _so = None
    _prefix = None
    _libs: dict = {}

    def __init__(self, **props):
        key = type(self).__name__
        if key not in _Component._libs:
            if not os.path.isfile(self._so):
                raise FileNotFoundError(f"{self._so} not built (run oracle.build() / make -C oracle)")
            lib = C.CDLL(self._so)
            _Component._libs[key] = _bind(lib, self._prefix)
        self._f = _Component._libs[key]
        self._h = self._f("create")()
        self.configure(**props)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._f("destroy")(self._h)
                self._h = None
        except Exception:
            pass

    # -- property surface (psk_soft.prf.xml:23-60) ------------------------------------
    def configure(self, **props):
        for k, v in props.items():
            if k not in PROPS:
                raise KeyError(k)
            rc = self._f("configure")(self._h, k.encode(), float(v))
            assert rc == 0

    def query(self, name):
        return self._f("query")(self._h, name.encode())

    # -- one BULKIO packet -------------------------------------------------------------
    def push(self, iq, xdelta=0.01, mode=1, flushed=False, sri_changed=False):
        """iq: complex64 array (or float32 interleaved). Returns dict of the four out-port payloads."""
        a = np.ascontiguousarray(iq)
        if a.dtype == np.complex64:
            a = a.view(np.float32)
        a = np.ascontiguousarray(a, dtype=np.float32)
        rc = self._f("push")(self._h, a.ctypes.data, a.size, float(xdelta), int(mode), int(flushed), int(sri_changed))
        out = self._drain()
        out["rc"] = rc
        return out

    def _drain(self):
        res = {}
        for port, name, dt in ((0, "soft", np.float32), (1, "bits", np.int16), (2, "phase", np.float32), (3, "sidx", np.int16)):
            n = self._f("read")(self._h, port, None, 0)
            buf = np.empty(n, dtype=dt)
            if n:
                self._f("read")(self._h, port, buf.ctypes.data, n)
            res[name] = buf
        res["soft"] = res["soft"].view(np.complex64)
        return res

    def sri(self, port):
        return dict(count=self._f("sri_count")(self._h, port), xdelta=self._f("sri_xdelta")(self._h, port),
                    mode=self._f("sri_mode")(self._h, port), packets=self._f("packet_count")(self._h, port))

    # -- whole stream ------------------------------------------------------------------
    def demod(self, iq, packet_len=64000, xdelta=0.01, keep=True):
        a = np.ascontiguousarray(iq)
        if a.dtype == np.complex64:
            a = a.view(np.float32)
        a = np.ascontiguousarray(a, dtype=np.float32)
        n = a.size // 2
        S = int(self.query("samplesPerBaud"))
        cap = n // max(S, 1) + 2
        nb = C.c_size_t(0)
        if keep:
            soft = np.empty(cap, np.complex64); bits = np.empty(cap * 3, np.int16)
            phase = np.empty(cap, np.float32); sidx = np.empty(cap, np.int16)
            k = self._f("demod")(self._h, a.ctypes.data, n, int(packet_len), float(xdelta),
                                 soft.ctypes.data, bits.ctypes.data, phase.ctypes.data, sidx.ctypes.data,
                                 cap, cap * 3, C.byref(nb))
            ns = k if S > 1 else 0
            return dict(soft=soft[:k].copy(), bits=bits[:nb.value].copy(), phase=phase[:k].copy(), sidx=sidx[:ns].copy())
        k = self._f("demod")(self._h, a.ctypes.data, n, int(packet_len), float(xdelta), None, None, None, None, 0, 0, C.byref(nb))
        return dict(n_symbols=k, n_bits=nb.value)


class RefComponent(_Component):
    """The unmodified reference psk_soft_i behind stub framework headers."""
    _so = REF_SO
    _prefix = "ref_"


class OracleComponent(_Component):
    """The plain-C restatement (oracle/psk_oracle.c)."""
    _so = ORC_SO
    _prefix = "orc_"
