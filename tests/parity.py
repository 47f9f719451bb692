"""Comparison helpers shared by the parity tests.

Bar (BASELINE.json north_star): sampleIndex and bits BIT-EXACT; phase and soft within
1e-4 * max(1, |ref|).  With differential decoding the first symbol is relative to the
default-constructed `last` (inf/NaN, reference: cpp/psk_soft.cpp:488, cpp/psk_soft.h:70) and is
skipped for the float comparison exactly as the reference's own test does
(tests/test_psk_soft.py:199-202); its bits are still compared.
"""
import numpy as np

FLOAT_RTOL = 1e-4


def float_close(a, b, rtol=FLOAT_RTOL):
    with np.errstate(invalid="ignore", over="ignore"):
        return _float_close(a, b, rtol)


def _float_close(a, b, rtol):
    a = np.asarray(a); b = np.asarray(b)
    if np.iscomplexobj(a):
        mag = np.maximum(1.0, np.abs(b))
        err = np.abs(a - b)
    else:
        mag = np.maximum(1.0, np.abs(b))
        err = np.abs(a - b)
    bad = ~(err <= rtol * mag)
    # non-finite values (division by a zero sample in differential mode, cpp/psk_soft.cpp:488) must agree
    # component by component: NaN with NaN, +/-inf with the same infinity
    def same_nonfinite(x, y):
        return (np.isnan(x) & np.isnan(y)) | (np.isinf(x) & np.isinf(y) & (np.sign(x) == np.sign(y)))
    if np.iscomplexobj(a):
        fin = np.isfinite(b.real) & np.isfinite(b.imag)
        agree = (same_nonfinite(a.real, b.real) | (np.isfinite(b.real) & (np.abs(a.real - b.real) <= rtol * mag))) & \
                (same_nonfinite(a.imag, b.imag) | (np.isfinite(b.imag) & (np.abs(a.imag - b.imag) <= rtol * mag)))
    else:
        fin = np.isfinite(b)
        agree = same_nonfinite(a, b)
    with np.errstate(invalid="ignore"):
        bad = np.where(fin, bad, ~agree)
        rel = np.where(fin, err / mag, 0.0)
    return bad, rel


def assert_parity(got, ref, differential=False, tag="", check_first_bits=True):
    assert len(got["sidx"]) == len(ref["sidx"]), f"{tag}: symbol count {len(got['sidx'])} != {len(ref['sidx'])}"
    assert len(got["bits"]) == len(ref["bits"]), f"{tag}: bit count {len(got['bits'])} != {len(ref['bits'])}"
    assert np.array_equal(got["sidx"], ref["sidx"]), f"{tag}: sampleIndex differs at {np.nonzero(got['sidx'] != ref['sidx'])[0][:8]}"
    gb, rb = got["bits"], ref["bits"]
    if differential and not check_first_bits and len(ref["sidx"]):
        bpb = len(rb) // len(ref["sidx"])
        gb, rb = gb[bpb:], rb[bpb:]
    assert np.array_equal(gb, rb), f"{tag}: bits differ at {np.nonzero(gb != rb)[0][:8]} of {len(rb)}"
    if "hard" in got and len(ref["sidx"]) and len(ref["bits"]):
        # the additional packed output: one byte per symbol = its bits, LSB first (checked against the REFERENCE's bits)
        bpb = len(ref["bits"]) // len(ref["sidx"])
        packed = (ref["bits"].reshape(-1, bpb).astype(np.uint8) << np.arange(bpb, dtype=np.uint8)).sum(axis=1).astype(np.uint8)
        gh = got["hard"]
        if "hard_mask" in got:
            gh = np.where(got["hard_mask"], packed, gh)
        s1 = 1 if (differential and not check_first_bits) else 0
        assert np.array_equal(gh[s1:], packed[s1:]), f"{tag}: packed hard symbols differ at {np.nonzero(gh != packed)[0][:8]}"
    bad, rel = float_close(got["phase"], ref["phase"])
    assert not bad.any(), f"{tag}: phase differs at {np.nonzero(bad)[0][:8]} max rel {np.nanmax(rel):.3e}"
    s0 = 1 if differential else 0
    bad, rel = float_close(got["soft"][s0:], ref["soft"][s0:])
    assert not bad.any(), f"{tag}: soft differs at {np.nonzero(bad)[0][:8] + s0} max rel {np.nanmax(rel):.3e}"
    return dict(max_rel_phase=float(np.nanmax(float_close(got["phase"], ref["phase"])[1])) if len(ref["phase"]) else 0.0,
                max_rel_soft=float(np.nanmax(rel)) if len(rel) else 0.0)


def checkers(oracle):
    """the CPU checkers to compare against: the C port always, the unmodified reference build (oracle/_ref) where it
    exists on this box -- (name, component class) pairs"""
    out = [("port", oracle.OracleComponent)]
    if oracle.have_ref():
        out.append(("reference", oracle.RefComponent))
    return out
