"""Synthetic PSK streams for the parity tests.

``gen_psk_reference`` is a Python-3 port of the generator in the reference's own test
(reference: tests/test_psk_soft.py:98-117, seed at :41).  Python 2's ``random.choice`` is
``seq[int(random() * len(seq))]``; ``random.seed(100)`` initialises the Mersenne Twister the
same way in 2 and 3, so this reproduces the reference's exact streams.

``gen_shaped`` is the harder fixture of SURVEY.md section 8d: pulse-shaped envelope (so the
timing argmax is non-degenerate), carrier offset, phase-noise random walk and complex AWGN.
"""
from __future__ import annotations

import math
import random

import numpy as np

REFERENCE_CASE_ORDER = (  # unittest runs the six reference cases alphabetically on one RNG stream
    ("DiffDecode8PSK", 8, True), ("DiffDecodeBPSK", 2, True), ("DiffDecodeQPSK", 4, True),
    ("NonDiffDecode8PSK", 8, False), ("NonDiffDecodeBPSK", 2, False), ("NonDiffDecodeQPSK", 4, False),
)


def gen_psk_reference(rng: random.Random, num_symbols: int, samp_per_baud: int = 8, num_syms: int = 4,
                      differential: bool = False):
    """reference: tests/test_psk_soft.py:98-117 (genPsk).  Returns (complex64 samples, list of input symbols)."""
    syms = list(range(num_syms))
    phase = [2 * math.pi * x / num_syms for x in syms]
    cx = [complex(math.cos(x), math.sin(x)) for x in phase]
    out = []
    input_symbols = []
    last = 1
    for _ in range(num_symbols):
        x = syms[int(rng.random() * len(syms))]      # python-2 random.choice
        x_cx = cx[x]
        input_symbols.append(x_cx)
        if differential:
            val = x_cx * last
            last = val
        else:
            val = x_cx
        for _ in range(samp_per_baud):
            out.append(val + .0001 * rng.random())
    return np.asarray(out, dtype=np.complex128).astype(np.complex64), input_symbols


def reference_cases(num_symbols: int = 1000, samp_per_baud: int = 8):
    """The six streams of the reference test module, generated in its execution order from seed 100."""
    rng = random.Random(100)
    cases = {}
    for name, m, diff in REFERENCE_CASE_ORDER:
        data, syms = gen_psk_reference(rng, num_symbols, samp_per_baud, m, diff)
        cases[name] = dict(iq=data, syms=np.asarray(syms, dtype=np.complex128), M=m, differential=diff,
                           samplesPerBaud=samp_per_baud)
    return cases


def gen_shaped(n_samples: int, sps: int, M: int, seed: int, sigma: float = 0.02, freq: float = 0.0,
               phase0: float = 0.0, pn_sigma: float = 0.0, timing_shift: int = 0, amp: float = 1.0):
    """unit-amplitude M-PSK, envelope 0.6+0.4*sin(pi*(p+0.5)/sps), carrier offset `freq`
    (cycles/sample), phase-noise walk (rad/sample sigma), complex AWGN sigma per dimension.
    Returns complex64."""
    rs = np.random.RandomState(seed)
    n_sym = (n_samples + timing_shift) // sps + 2
    sym = rs.randint(0, M, size=n_sym)
    n = np.arange(n_samples)
    pos = n + timing_shift
    k = pos // sps
    p = pos % sps
    env = 0.6 + 0.4 * np.sin(np.pi * (p + 0.5) / sps)
    ph = 2 * np.pi * sym[k] / M + 2 * np.pi * freq * n + phase0
    if pn_sigma > 0:
        ph = ph + np.cumsum(rs.normal(0.0, pn_sigma, size=n_samples))
    x = amp * env * np.exp(1j * ph)
    x = x + sigma * (rs.normal(size=n_samples) + 1j * rs.normal(size=n_samples))
    return x.astype(np.complex64)
