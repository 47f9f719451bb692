"""Python mirror of the reference component's interface over the C ABI.

``Bank``     n_channels independent demodulators on one GPU (one pskd handle).
``PskSoft``  a single component: same property names as psk_soft.prf.xml:23-60, ``push()`` is
             one BULKIO packet through serviceFunction (reference: cpp/psk_soft.cpp:346-618),
             returning what the four out-ports would carry.

All compute happens in libpskd.so; numpy is only used to hold host buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding as B

PROP_NAMES = ("samplesPerBaud", "numAvg", "constelationSize", "phaseAvg", "differentialDecoding", "resetState")


def default_props() -> dict:
    p = B.Props()
    B.load().pskd_default_props(C.byref(p))
    return {n: int(getattr(p, n)) for n in PROP_NAMES}


def _to_props(d: dict) -> B.Props:
    base = default_props()
    for k, v in d.items():
        if k not in PROP_NAMES:
            raise KeyError(k)
        base[k] = int(v)
    return B.Props(**base)


class Bank:
    """A bank of independent channels on one GPU (reference: one psk_soft_i per channel)."""

    def __init__(self, n_channels: int, props=None, device: int = 0):
        self.lib = B.load()
        self.n_channels = int(n_channels)
        self.device = int(device)
        if props is None:
            arr = None
        else:
            if isinstance(props, dict):
                props = [props] * self.n_channels
            assert len(props) == self.n_channels
            arr = (B.Props * self.n_channels)(*[_to_props(p) for p in props])
        self._h = C.c_void_p()
        rc = self.lib.pskd_create(C.byref(self._h), self.device, self.n_channels, arr)
        self._check(rc)

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise B.PskdError(rc, (self.lib.pskd_last_error() or b"").decode())
        return rc

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.pskd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self.lib.pskd_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.pskd_launch_count(self._h))

    def sync(self):
        self._check(self.lib.pskd_sync(self._h))

    # -- property surface ------------------------------------------------------------------
    def set_props(self, ch: int = -1, **props):
        """Change the named properties of channel `ch` (-1: of every channel, each keeping its other properties)."""
        for c in (range(self.n_channels) if ch < 0 else (ch,)):
            cur = self.get_props(c)
            cur.update({k: int(v) for k, v in props.items()})
            self._check(self.lib.pskd_set_props(self._h, int(c), C.byref(_to_props(cur))))

    def get_props(self, ch: int = 0) -> dict:
        p = B.Props()
        self._check(self.lib.pskd_get_props(self._h, int(ch), C.byref(p)))
        return {n: int(getattr(p, n)) for n in PROP_NAMES}

    def max_symbols(self, n_complex: int, ch: int = 0) -> int:
        return int(self.lib.pskd_max_symbols(self._h, ch, int(n_complex)))

    def sri(self, ch: int = 0) -> dict:
        s = B.SriOut()
        self._check(self.lib.pskd_get_sri(self._h, ch, C.byref(s)))
        return {n: getattr(s, n) for n, _ in B.SriOut._fields_}

    def stats(self) -> dict:
        s = B.Stats()
        self._check(self.lib.pskd_get_stats(self._h, C.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in B.Stats._fields_}

    def profile_enable(self, on: bool = True):
        self._check(self.lib.pskd_profile_enable(self._h, int(on)))

    def profile_read(self, reset: bool = True) -> dict:
        """{kernel name: (total ms, launches, algorithmic bytes)} measured with CUDA events on the bank's stream."""
        arr = (B.KernelTime * 24)()
        n = C.c_int(0)
        self._check(self.lib.pskd_profile_read(self._h, arr, 24, C.byref(n), int(reset)))
        return {arr[i].name.decode(): (arr[i].ms_total, int(arr[i].launches), float(arr[i].alg_bytes)) for i in range(min(n.value, 24))}

    # -- checkpoint / resume -----------------------------------------------------------------
    def export_state(self) -> bytes:
        """the bank's carried state (cpp/psk_soft.h:66-86 of every channel) as one blob"""
        n = int(self.lib.pskd_state_size(self._h))
        buf = C.create_string_buffer(n)
        w = C.c_size_t(0)
        self._check(self.lib.pskd_state_export(self._h, buf, n, C.byref(w)))
        return buf.raw[:w.value]

    def import_state(self, blob: bytes):
        buf = C.create_string_buffer(blob, len(blob))
        self._check(self.lib.pskd_state_import(self._h, buf, len(blob)))

    # -- the hot path ----------------------------------------------------------------------
    def process_raw(self, iq_ptr, iq_stride, n_complex, soft_ptr, bits_ptr, phase_ptr, sidx_ptr, sym_stride,
                    bits_stride, xdelta=0.01, packet_len=64000, flags=0, sri_mode=1, counts=True, hard_ptr=None):
        """Direct call of pskd_process with raw pointers (device pointers unless FLAG_HOST_BUFFERS)."""
        inp = B.Input()
        inp.iq = iq_ptr
        inp.iq_stride = int(iq_stride)
        if np.isscalar(n_complex):
            inp.n_complex = None
            inp.n_complex_all = int(n_complex)
        else:
            self._ncx = (C.c_size_t * self.n_channels)(*[int(v) for v in n_complex])
            inp.n_complex = self._ncx
            inp.n_complex_all = 0
        inp.sri_xdelta = float(xdelta)
        inp.sri_mode = int(sri_mode)
        inp.packet_len = int(packet_len)
        inp.flags = int(flags)
        out = B.Output()
        out.soft, out.bits, out.phase, out.sample_index = soft_ptr, bits_ptr, phase_ptr, sidx_ptr
        out.hard = hard_ptr
        out.sym_stride, out.bits_stride = int(sym_stride), int(bits_stride)
        if counts:
            self._nsym = (C.c_size_t * self.n_channels)()
            self._nbits = (C.c_size_t * self.n_channels)()
            out.n_symbols, out.n_bits = self._nsym, self._nbits
        rc = self._check(self.lib.pskd_process(self._h, C.byref(inp), C.byref(out)))
        if counts:
            return rc, np.frombuffer(self._nsym, dtype=np.uintp).copy(), np.frombuffer(self._nbits, dtype=np.uintp).copy()
        return rc, None, None

    def process_host(self, iq, n_complex=None, xdelta=0.01, packet_len=64000, flushed=False, sri_mode=1):
        """iq: complex64 [n_channels, n] (or [n] for one channel) in HOST memory.
        Returns a list (one dict per channel) of soft/bits/phase/sidx numpy arrays."""
        a = np.ascontiguousarray(iq, dtype=np.complex64)
        if a.ndim == 1:
            a = a[None, :]
        assert a.shape[0] == self.n_channels
        n = a.shape[1]
        ncx = n if n_complex is None else n_complex
        nmax = n if n_complex is None else int(max(n_complex))
        cap = max(self.max_symbols(nmax, ch) for ch in (range(self.n_channels) if self.n_channels <= 64 else (0,))) + 8
        if self.n_channels > 64:
            cap = max(cap, nmax // 2 + 8)
        soft = np.zeros((self.n_channels, cap), np.complex64)
        phase = np.zeros((self.n_channels, cap), np.float32)
        sidx = np.zeros((self.n_channels, cap), np.int16)
        bits = np.zeros((self.n_channels, 3 * cap), np.int16)
        hard = np.zeros((self.n_channels, cap), np.uint8)      # the additional packed hard-symbol output
        flags = B.FLAG_HOST_BUFFERS | (B.FLAG_QUEUE_FLUSHED if flushed else 0)
        rc, ns, nb = self.process_raw(a.ctypes.data if a.size else None, n, ncx, soft.ctypes.data, bits.ctypes.data,
                                      phase.ctypes.data, sidx.ctypes.data, cap, 3 * cap, xdelta, packet_len, flags, sri_mode,
                                      hard_ptr=hard.ctypes.data)
        res = []
        for c in range(self.n_channels):
            k, b = int(ns[c]), int(nb[c])
            res.append(dict(soft=soft[c, :k].copy(), bits=bits[c, :b].copy(), phase=phase[c, :k].copy(),
                            sidx=sidx[c, :k].copy(), hard=hard[c, :k].copy(), rc=rc))
        return res


class PskSoft:
    """One component instance with the reference's property names; push() = one BULKIO packet."""

    def __init__(self, device: int = 0, **props):
        object.__setattr__(self, "_bank", Bank(1, [props] if props else None, device))

    def __getattr__(self, name):
        if name in PROP_NAMES:
            return self._bank.get_props(0)[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in PROP_NAMES:
            self._bank.set_props(0, **{name: value})
        else:
            object.__setattr__(self, name, value)

    def configure(self, **props):
        self._bank.set_props(0, **props)

    def push(self, iq, xdelta=0.01, mode=1, flushed=False):
        """One packet through the demod core (reference: cpp/psk_soft.cpp:349-618)."""
        return self._bank.process_host(np.asarray(iq, np.complex64), xdelta=xdelta, packet_len=0,
                                       flushed=flushed, sri_mode=mode)[0]

    def demod(self, iq, packet_len=64000, xdelta=0.01):
        """A whole stream, processed as if delivered in packets of packet_len complex samples."""
        return self._bank.process_host(np.asarray(iq, np.complex64), xdelta=xdelta, packet_len=packet_len)[0]

    def sri(self):
        return self._bank.sri(0)

    def stats(self):
        return self._bank.stats()


def synth_fill(iq_dev_ptr: int, iq_stride: int, ch0: int, n_channels: int, n_complex: int, *, seed: int,
               samplesPerBaud: int, constelationSize: int, sigma: float = 0.02, freq_max: float = 2e-5,
               pn_sigma: float = 0.0, shaped: bool = True, device: int = 0, stream: int = 0, period: int = 0):
    """Fill a device buffer with a synthetic PSK channel bank (see include/pskd.h: pskd_synth_fill).
    period > 0: carrier offsets quantised so that a buffer of `period` samples can be replayed as one continuous stream."""
    lib = B.load()
    cfg = B.Synth(seed, samplesPerBaud, constelationSize, sigma, freq_max, pn_sigma, 1.0 if shaped else 0.0, int(period))
    rc = lib.pskd_synth_fill(device, iq_dev_ptr, int(iq_stride), int(ch0), int(n_channels), int(n_complex),
                             C.byref(cfg), stream or None)
    if rc != 0:
        raise B.PskdError(rc, "pskd_synth_fill failed")
