"""CPU tests of the checker itself: the plain-C restatement (oracle/psk_oracle.c) against
(a) the committed golden vectors (generated from the unmodified reference build by
tests/golden/make_golden.py) and (b) -- where /root/reference exists -- the reference build
itself on fresh seeded inputs.  Bit-for-bit on all four ports (these are both x86 CPU code using
the same libm)."""
import glob
import os

import numpy as np
import pytest

import siggen
from parity import assert_parity

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
PROP_NAMES = ("samplesPerBaud", "numAvg", "constelationSize", "phaseAvg", "differentialDecoding")


def load_golden(path):
    z = np.load(path)
    props = {k: int(v) for k, v in zip(PROP_NAMES, z["props"])}
    ref = dict(soft=z["soft"], bits=z["bits"], phase=z["phase"], sidx=z["sidx"])
    return z["iq"], props, int(z["packet_len"]), float(z["xdelta"]), ref


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_golden_files_present():
    assert len(GOLDEN) >= 12


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_golden_bit_for_bit(path, oracle_built):
    iq, props, pkt, xd, ref = load_golden(path)
    got = oracle_built.OracleComponent(**props).demod(iq, packet_len=pkt, xdelta=xd)
    for k in ("sidx", "bits"):
        assert np.array_equal(got[k], ref[k]), k
    # same libm on the same image: floats are bit-identical too (NaN-safe comparison)
    for k in ("phase", "soft"):
        if not bits_equal(got[k], ref[k]):
            assert_parity(got, ref, differential=bool(props["differentialDecoding"]), tag=path)


CASES = [
    dict(S=8, M=4, A=100, P=50, D=0, pkt=64000, xd=0.01, sig=0.02, f=1e-5),
    dict(S=10, M=2, A=100, P=50, D=0, pkt=6400, xd=0.01, sig=0.05, f=1e-4, pn=0.002),
    dict(S=8, M=8, A=100, P=50, D=1, pkt=8000, xd=0.01, sig=0.02, f=2e-5),
    dict(S=8, M=8, A=37, P=20, D=0, pkt=1001, xd=1.0, sig=0.02, f=2e-5),
    dict(S=9, M=2, A=64, P=50, D=1, pkt=777, xd=0.01, sig=0.05, f=0.0),
    dict(S=8, M=4, A=100, P=50, D=0, pkt=64, xd=0.01, sig=0.02, f=1e-5),
    dict(S=8, M=3, A=10, P=5, D=0, pkt=500, xd=0.5, sig=0.02, f=1e-5),
    dict(S=1, M=4, A=0, P=5, D=0, pkt=500, xd=0.5, sig=0.02, f=1e-5),
    dict(S=8, M=8, A=100, P=50, D=0, pkt=16000, xd=0.01, sig=0.15, f=2e-5),
]


@pytest.mark.parametrize("t", CASES, ids=lambda t: f"S{t['S']}M{t['M']}A{t['A']}P{t['P']}D{t['D']}pkt{t['pkt']}")
def test_oracle_matches_reference_build(t, oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here); golden vectors pin the oracle instead")
    iq = siggen.gen_shaped(120000, t["S"], t["M"], seed=5, sigma=t["sig"], freq=t["f"], pn_sigma=t.get("pn", 0), timing_shift=3)
    props = dict(samplesPerBaud=t["S"], constelationSize=t["M"], numAvg=t["A"], phaseAvg=t["P"], differentialDecoding=t["D"])
    r = oracle_built.RefComponent(**props).demod(iq, packet_len=t["pkt"], xdelta=t["xd"])
    o = oracle_built.OracleComponent(**props).demod(iq, packet_len=t["pkt"], xdelta=t["xd"])
    for k in ("soft", "bits", "phase", "sidx"):
        assert bits_equal(r[k], o[k]), f"{k} differs"


def test_reconfiguration_sequence_matches_reference(oracle_built):
    """scripted property changes / resets between packets (reference: cpp/psk_soft.cpp:353-426, 638-651)"""
    if not oracle_built.have_ref():
        pytest.skip("needs oracle/_ref")
    iq = siggen.gen_shaped(200000, 8, 4, seed=9, sigma=0.03, freq=2e-5, timing_shift=2)
    comps = [oracle_built.RefComponent(samplesPerBaud=8, constelationSize=4), oracle_built.OracleComponent(samplesPerBaud=8, constelationSize=4)]
    script = [(0, 30000, {}), (30000, 60000, dict(phaseAvg=20)), (60000, 90000, dict(constelationSize=8)),
              (90000, 120000, dict(resetState=1)), (120000, 150000, dict(numAvg=150)), (150000, 200000, dict(differentialDecoding=1))]
    for a, b, ch in script:
        outs = []
        for c in comps:
            c.configure(**ch)
            outs.append(c.push(iq[a:b], xdelta=0.01))
        for k in ("soft", "bits", "phase", "sidx"):
            assert bits_equal(outs[0][k], outs[1][k]), f"{k} differs after {ch}"
        assert comps[0].sri(0) == comps[1].sri(0) and comps[0].sri(1)["count"] == comps[1].sri(1)["count"]


def test_stall_after_window_shrink_matches_reference(oracle_built):
    """numAvg*samplesPerBaud shrinks below the carried window: the reference consumes packets and emits nothing until the
    window grows past the deque (cpp/psk_soft.cpp:380-383, 457, 619-636); resetState only truncates the deque.  The C port
    must stall and recover exactly like the reference build (the GPU test runs the same script against the port)."""
    if not oracle_built.have_ref():
        pytest.skip("needs oracle/_ref")
    iq = siggen.gen_shaped(120000, 8, 8, seed=13, sigma=0.03, freq=2e-5, timing_shift=3)
    props = dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50)
    comps = [oracle_built.RefComponent(**props), oracle_built.OracleComponent(**props)]
    script = [(0, 20000, {}), (20000, 20300, dict(numAvg=50)), (20300, 20301, {}), (20301, 20500, dict(constelationSize=4)),
              (20500, 40000, dict(numAvg=200, constelationSize=8)), (40000, 60000, dict(numAvg=20)), (60000, 60700, dict(resetState=1)),
              (60700, 61000, dict(phaseAvg=30)), (61000, 90000, dict(numAvg=150)), (90000, 120000, {})]
    emitted = []
    for a, b, ch in script:
        outs = []
        for c in comps:
            c.configure(**ch)
            outs.append(c.push(iq[a:b], xdelta=0.01))
        for k in ("soft", "bits", "phase", "sidx"):
            assert bits_equal(outs[0][k], outs[1][k]), f"{k} differs after {ch}"
        assert comps[0].sri(0) == comps[1].sri(0)
        emitted.append(len(outs[0]["sidx"]))
    assert emitted[1:4] == [0, 0, 0] and emitted[4] > 0 and emitted[5:8] == [0, 0, 0] and emitted[8] > 0, emitted


@pytest.mark.parametrize("name", [c[0] for c in siggen.REFERENCE_CASE_ORDER])
def test_reference_own_assertions_hold(name, oracle_built):
    """The assertions of the reference's test module (tests/test_psk_soft.py:178-238): soft-decision
    error < 1e-3 against the transmitted symbols (differential: first output skipped, QPSK rotated by
    pi/4; coherent: best of the M admissible rotations)."""
    import math
    c = siggen.reference_cases()[name]
    M = c["M"]
    props = dict(samplesPerBaud=8, constelationSize=M, numAvg=100, differentialDecoding=int(c["differential"]))
    out = oracle_built.OracleComponent(**props).demod(c["iq"], packet_len=64000, xdelta=0.01)["soft"].astype(np.complex128)
    syms = c["syms"][:len(out)]
    assert len(out) == 901
    if c["differential"]:
        rot = np.exp(1j * math.pi / 4) if M == 4 else 1.0
        err = np.max(np.abs(out[1:] - rot * syms[1:]))
    else:
        thetas = {2: [0, math.pi], 4: [math.pi / 4 * k for k in (1, 3, 5, 7)], 8: [math.pi / 4 * k for k in range(8)]}[M]
        err = min(np.max(np.abs(np.exp(1j * th) * out[1:] - syms[1:])) for th in thetas)
    assert err < 1e-3
