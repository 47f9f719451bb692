"""GPU vs the committed golden vectors (outputs of the unmodified reference build), through the
C ABI.  Also scripted reconfiguration / reset sequences against the oracle."""
import glob
import os

import numpy as np
import pytest

import siggen
from parity import assert_parity

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
PROP_NAMES = ("samplesPerBaud", "numAvg", "constelationSize", "phaseAvg", "differentialDecoding")


@pytest.fixture(params=["fused", "staged", "tp", "legacy"], autouse=True)
def fused_mode(request, monkeypatch):
    """every test of this module runs through the fused kernel, the staged kernels (k_fzs_front +
    k_fzs_cb), the staged kernels with the time-parallel chain, and the legacy staged kernels
    (PSKD_FZS=0: k_front_t + k_chain_par + k_back_par, time-parallel chain on)"""
    monkeypatch.setenv("PSKD_FUSED", "1" if request.param == "fused" else "0")
    monkeypatch.setenv("PSKD_TP", "1" if request.param in ("tp", "legacy") else "0")
    monkeypatch.setenv("PSKD_FZS", "0" if request.param == "legacy" else "1")
    return request.param


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_gpu_matches_golden(path):
    import psk_soft_b200 as pk
    z = np.load(path)
    props = {k: int(v) for k, v in zip(PROP_NAMES, z["props"])}
    ref = dict(soft=z["soft"], bits=z["bits"], phase=z["phase"], sidx=z["sidx"])
    got = pk.PskSoft(**props).demod(z["iq"], packet_len=int(z["packet_len"]), xdelta=float(z["xdelta"]))
    assert_parity(got, ref, differential=bool(props["differentialDecoding"]), tag=os.path.basename(path))


def test_reconfiguration_script_matches_oracle(oracle_built):
    """property changes and resets between packets: phaseAvg / constelationSize listeners,
    resetState, queue flush, numAvg and samplesPerBaud growth, differential toggle
    (reference: cpp/psk_soft.cpp:353-426, 619-651)"""
    import psk_soft_b200 as pk
    iq = siggen.gen_shaped(330000, 8, 4, seed=9, sigma=0.03, freq=2e-5, timing_shift=2)
    orc = oracle_built.OracleComponent(samplesPerBaud=8, constelationSize=4)
    dev = pk.PskSoft(samplesPerBaud=8, constelationSize=4)
    script = [(0, 30000, {}, False), (30000, 60000, dict(phaseAvg=20), False), (60000, 90000, dict(constelationSize=8), False),
              (90000, 120000, dict(resetState=1), False), (120000, 150000, dict(numAvg=150), False),
              (150000, 200000, dict(differentialDecoding=1), False), (200000, 230000, {}, True),
              (230000, 280000, dict(samplesPerBaud=10, differentialDecoding=0), False), (280000, 330000, dict(phaseAvg=64), False)]
    for a, b, ch, flushed in script:
        orc.configure(**ch)
        dev.configure(**ch)
        ref = orc.push(iq[a:b], xdelta=0.01, flushed=flushed)
        got = dev.push(iq[a:b], xdelta=0.01, flushed=flushed)
        diff = bool(dev.differentialDecoding)
        # after toggling differential on, `last` is whatever was carried: compare every symbol's bits, skip only inf/NaN floats
        assert_parity(got, ref, differential=False if np.isfinite(ref["soft"]).all() else diff, tag=f"{a}:{b} {ch}")
        s_ref, s_dev = orc.sri(0), dev.sri()
        assert abs(s_dev["soft_xdelta"] - s_ref["xdelta"]) < 1e-15 and s_dev["soft_mode"] == s_ref["mode"]
        assert s_dev["sri_pushes"] == s_ref["count"]


def test_sri_metadata_matches_oracle(oracle_built):
    """out-port SRIs (cpp/psk_soft.cpp:393-405): soft xdelta*S mode 1, phase mode 0, bits xdelta*S/b"""
    import psk_soft_b200 as pk
    iq = siggen.gen_shaped(50000, 10, 8, seed=1)
    props = dict(samplesPerBaud=10, constelationSize=8)
    orc = oracle_built.OracleComponent(**props)
    dev = pk.PskSoft(**props)
    orc.demod(iq, packet_len=6400, xdelta=0.004)
    dev.demod(iq, packet_len=6400, xdelta=0.004)
    s = dev.sri()
    assert s["soft_xdelta"] == orc.sri(0)["xdelta"] and s["soft_mode"] == 1
    assert s["phase_xdelta"] == orc.sri(2)["xdelta"] and s["phase_mode"] == 0
    assert s["bits_xdelta"] == orc.sri(1)["xdelta"] and s["bits_mode"] == 0
    assert s["sri_pushes"] == orc.sri(0)["count"] == 8


def test_unsupported_and_error_paths():
    import psk_soft_b200 as pk
    with pytest.raises(pk.PskdError):
        pk.PskSoft(samplesPerBaud=1)                  # reference's sps==1 branch: not on the GPU path
    with pytest.raises(pk.PskdError) as e:
        pk.PskSoft(samplesPerBaud=8, numAvg=0)        # never emits in the reference
    assert e.value.code == -4


def test_stall_after_window_shrink_matches_oracle(oracle_built):
    """numAvg*samplesPerBaud shrinks below the carried window: the reference keeps consuming packets and emits
    NOTHING (`samples.size()==numDataPts`, cpp/psk_soft.cpp:457, cannot become true while the deque only grows) until
    the window length grows past the deque size; resetState only truncates the deque to its oldest samples
    (resyncEnergy, :619-636) and the stall goes on.  The C ABI emulates exactly that (no error)."""
    import psk_soft_b200 as pk
    iq = siggen.gen_shaped(120000, 8, 8, seed=13, sigma=0.03, freq=2e-5, timing_shift=3)
    props = dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50)
    orc = oracle_built.OracleComponent(**props)
    dev = pk.PskSoft(**props)
    script = [(0, 20000, {}), (20000, 20300, dict(numAvg=50)),                 # shrink -> stall (deque 799+300)
              (20300, 20301, {}), (20301, 20500, dict(constelationSize=4)),    # still stalled; a pending reset runs its prologue
              (20500, 40000, dict(numAvg=200, constelationSize=8)),            # 1600 > deque size: recovers, window = the grown deque
              (40000, 60000, dict(numAvg=20)), (60000, 60700, dict(resetState=1)),   # stall; reset truncates, stall goes on
              (60700, 61000, dict(phaseAvg=30)), (61000, 90000, dict(numAvg=150)),   # 1200 > 160+700+300: recovers
              (90000, 120000, {})]
    emitted = []
    for a, b, ch in script:
        orc.configure(**ch)
        dev.configure(**ch)
        ref = orc.push(iq[a:b], xdelta=0.01)
        got = dev.push(iq[a:b], xdelta=0.01)
        emitted.append(len(ref["sidx"]))
        assert_parity(got, ref, tag=f"{a}:{b} {ch}")
        s_ref, s_dev = orc.sri(0), dev.sri()
        assert s_dev["sri_pushes"] == s_ref["count"], (a, b, s_dev, s_ref)
    assert emitted[1] == emitted[2] == emitted[3] == 0 and emitted[4] > 0, emitted
    assert emitted[5] == emitted[6] == emitted[7] == 0 and emitted[8] > 0, emitted


def test_state_import_rejects_corrupt_blobs():
    """pskd_state_import validates the blob against its own size and the bank before touching the bank"""
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50)
    bank = pk.Bank(2, props)
    iq = np.stack([siggen.gen_shaped(30000, 8, 8, seed=60 + c) for c in range(2)])
    first = bank.process_host(iq[:, :15000].copy(), xdelta=0.01, packet_len=8000)
    blob = bytearray(bank.export_state())
    import struct
    hdr = struct.Struct("<IIiiQ")
    magic, ver, nch, ring_cap, total = hdr.unpack_from(blob, 0)
    assert total == len(blob)

    def rejected(mut, code=-1):
        with pytest.raises(pk.PskdError) as e:
            bank.import_state(bytes(mut))
        assert e.value.code == code, e.value

    rejected(blob[:len(blob) - 8])                                    # truncated
    rejected(blob[:hdr.size + 10])
    bad = bytearray(blob); hdr.pack_into(bad, 0, magic, ver, nch, 1 << 30, total); rejected(bad)   # absurd ring
    bad = bytearray(blob); hdr.pack_into(bad, 0, magic, ver, nch, ring_cap, total - 4); rejected(bad)
    bad = bytearray(blob); bad[hdr.size + 0:hdr.size + 2] = (1).to_bytes(2, "little"); rejected(bad, -4)   # samplesPerBaud = 1 in channel 0's properties
    # corrupt LinearFit head / n of channel 0 (offsets inside the first StateChan are implementation details: flip every
    # int32 of the device part that currently equals phaseAvg or the ring head and expect either a rejection or parity)
    second = bank.process_host(iq[:, 15000:].copy(), xdelta=0.01, packet_len=8000)
    ref = pk.Bank(2, props)
    r1 = ref.process_host(iq[:, :15000].copy(), xdelta=0.01, packet_len=8000)
    ref.import_state(bytes(blob))                                      # a valid blob still imports ...
    r2 = ref.process_host(iq[:, 15000:].copy(), xdelta=0.01, packet_len=8000)
    for c in range(2):
        for k in ("sidx", "bits", "phase", "soft"):
            assert np.array_equal(second[c][k], r2[c][k], equal_nan=True), (c, k)   # ... and the rejected ones left the bank untouched


def test_cpp_host_mirror_demo_runs():
    """the C++ host-side mirror (psk_soft_b200/host/psk_soft_gpu.hpp) driving serviceFunction-shaped calls"""
    import subprocess
    from psk_soft_b200 import _build
    exe = os.path.join(_build.LIB_DIR, "demo_component")
    if not os.path.isfile(exe):
        _build.build_host_demo()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "total symbols 7901" in res.stdout


def test_checkpoint_resume_matches_uninterrupted(oracle_built):
    """pskd_state_export -> new bank -> pskd_state_import continues the streams exactly where the
    first bank stopped (carried state = the members at cpp/psk_soft.h:66-86 of every channel)."""
    import psk_soft_b200 as pk
    props = [dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50),
             dict(samplesPerBaud=10, constelationSize=4, numAvg=50, phaseAvg=25, differentialDecoding=1),
             dict(samplesPerBaud=9, constelationSize=2, numAvg=64, phaseAvg=100)]
    n, cut = 90000, 41234
    iq = np.zeros((3, n), np.complex64)
    for c, p in enumerate(props):
        iq[c] = siggen.gen_shaped(n, p["samplesPerBaud"], p["constelationSize"], seed=40 + c, sigma=0.03, freq=2e-5, timing_shift=c)
    whole = pk.Bank(3, props)
    w1 = whole.process_host(iq[:, :cut].copy(), xdelta=0.01, packet_len=8000)
    w2 = whole.process_host(iq[:, cut:].copy(), xdelta=0.01, packet_len=8000)
    first = pk.Bank(3, props)
    f1 = first.process_host(iq[:, :cut].copy(), xdelta=0.01, packet_len=8000)
    blob = first.export_state()
    first.close()
    second = pk.Bank(3)                       # default properties: everything comes from the blob
    second.import_state(blob)
    assert second.get_props(1)["differentialDecoding"] == 1
    f2 = second.process_host(iq[:, cut:].copy(), xdelta=0.01, packet_len=8000)
    for c in range(3):
        for k in ("sidx", "bits", "phase", "soft"):
            assert np.array_equal(w1[c][k], f1[c][k], equal_nan=(k in ("phase", "soft"))), (c, k)
            assert np.array_equal(w2[c][k], f2[c][k], equal_nan=(k in ("phase", "soft"))), (c, k)
        ref = oracle_built.OracleComponent(**props[c])
        r1 = ref.demod(iq[c, :cut], packet_len=8000, xdelta=0.01)
        r2 = ref.demod(iq[c, cut:], packet_len=8000, xdelta=0.01)
        d = bool(props[c].get("differentialDecoding", 0))
        assert_parity(f1[c], r1, differential=d, tag=f"ch{c} before the checkpoint")
        assert_parity(f2[c], r2, differential=False if np.isfinite(r2["soft"]).all() else d, tag=f"ch{c} after the restore")
    with pytest.raises(pk.binding.PskdError):
        pk.Bank(2).import_state(blob)         # channel count mismatch is an argument error


def test_tiny_and_ragged_calls(oracle_built):
    """calls that emit 0, 1, 31, 32, 33 ... symbols and packets shorter than one symbol: the fused
    kernel's partial chunks / partial chain blocks / empty packets against the oracle"""
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=8, constelationSize=8, numAvg=20, phaseAvg=10)
    iq = siggen.gen_shaped(40000, 8, 8, seed=77, sigma=0.03, freq=3e-5, timing_shift=4)
    orc = oracle_built.OracleComponent(**props)
    dev = pk.PskSoft(**props)
    cuts = [0, 5, 150, 159, 160, 168, 168 + 8 * 31, 168 + 8 * 63, 168 + 8 * 96, 2000, 2001, 2003, 9000, 9007, 30000, 40000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ref = orc.push(iq[a:b], xdelta=0.01)
        got = dev.push(iq[a:b], xdelta=0.01)
        assert_parity(got, ref, tag=f"call {a}:{b}")
    # the same stream in one call, cut into 3-sample packets (most packets emit nothing)
    ref = oracle_built.OracleComponent(**props).demod(iq[:6000], packet_len=3, xdelta=0.01)
    got = pk.PskSoft(**props).demod(iq[:6000], packet_len=3, xdelta=0.01)
    assert_parity(got, ref, tag="3-sample packets")


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("PSKD_FUZZ_SEEDS", "12")))))
def test_randomized_reconfiguration_scripts(seed, oracle_built):
    """random sequences of packets with property changes in between (any constellation / phaseAvg / differential
    toggle, resetState, queue flush, SRI rate changes, window growth) against the oracle, state carried across every
    call (reference: cpp/psk_soft.cpp:353-426, 619-651).  Window shrinks stall the component (:457 never true again until
    the window grows past the deque): emulated, both sides emit nothing."""
    import psk_soft_b200 as pk
    rs = np.random.RandomState(7000 + seed)
    S, A = int(rs.choice([8, 9, 10, 16])), int(rs.choice([3, 20, 64, 100]))
    props = dict(samplesPerBaud=S, numAvg=A, constelationSize=int(rs.choice([2, 4, 8])), phaseAvg=int(rs.choice([2, 10, 50, 64])),
                 differentialDecoding=int(rs.randint(0, 2)))
    orc = oracle_built.OracleComponent(**props)
    dev = pk.PskSoft(**props)
    xdelta = 0.01
    for step in range(14):
        ch = {}
        r = rs.rand()
        if r < 0.15:
            ch["constelationSize"] = int(rs.choice([2, 4, 8]))
        elif r < 0.30:
            ch["phaseAvg"] = int(rs.choice([2, 5, 25, 50, 100, 128]))
        elif r < 0.42:
            ch["differentialDecoding"] = int(rs.randint(0, 2))
        elif r < 0.50:
            ch["resetState"] = 1
        elif r < 0.62:                                         # window growth; one change in four may shrink it (stall)
            S2, A2 = int(rs.choice([8, 9, 10, 16])), int(rs.choice([3, 20, 64, 100, 150]))
            if S2 * A2 >= S * A or rs.rand() < 0.25:
                if S2 != S: ch["samplesPerBaud"] = S2
                if A2 != A: ch["numAvg"] = A2
                S, A = S2, A2
        if rs.rand() < 0.15:
            xdelta = float(rs.choice([0.01, 0.02, 1.0, 0.5]))
        flushed = bool(rs.rand() < 0.1)
        n = int(rs.choice([1, 7, S * A // 2 + 1, 2000, 9000, 30000, 70000]))
        M = int(dev.constelationSize if "constelationSize" not in ch else ch["constelationSize"])
        iq = siggen.gen_shaped(n, S, max(M, 2), seed=int(rs.randint(1 << 30)), sigma=0.03, freq=float(rs.uniform(-3e-5, 3e-5)),
                               phase0=float(rs.uniform(0, 6.28)), timing_shift=int(rs.randint(0, S)))
        orc.configure(**ch)
        dev.configure(**ch)
        ref = orc.push(iq, xdelta=xdelta, flushed=flushed)
        got = dev.push(iq, xdelta=xdelta, flushed=flushed)
        diff = bool(dev.differentialDecoding)
        assert_parity(got, ref, differential=False if np.isfinite(ref["soft"]).all() else diff,
                      tag=f"seed {seed} step {step} n {n} {ch} xdelta {xdelta} flushed {flushed}")


def test_host_buffer_slab_ring_and_async_calls(oracle_built, monkeypatch):
    """host-buffer calls cut into MANY slabs (PSKD_SLAB_MB=1: one or two channels per slab, more slabs than ring slots, each slab
    with its own time-parallel plan), synchronous and pipelined (PSKD_FLAG_NO_SYNC + pskd_sync, two calls in flight with their
    own output buffers) -- every channel of every call against the oracle"""
    import psk_soft_b200 as pk
    from psk_soft_b200 import binding as B
    monkeypatch.setenv("PSKD_SLAB_MB", "1")
    rs = np.random.RandomState(12)
    nch, n = 20, 64000
    props, iqs = [], []
    for c in range(nch):
        S = int(rs.choice([8, 10])); M = int(rs.choice([2, 4, 8])); D = int(rs.randint(0, 2))
        props.append(dict(samplesPerBaud=S, constelationSize=M, numAvg=int(rs.choice([50, 100])), phaseAvg=int(rs.choice([25, 50])), differentialDecoding=D))
        iqs.append(siggen.gen_shaped(2 * n, S, M, seed=300 + c, sigma=0.03, freq=float(rs.uniform(-2e-5, 2e-5)), timing_shift=c % S))
    iqs = np.stack(iqs)
    # (a) synchronous calls
    bank = pk.Bank(nch, props)
    orcs = [oracle_built.OracleComponent(**p) for p in props]
    for a, b in ((0, n), (n, 2 * n)):
        got = bank.process_host(iqs[:, a:b].copy(), xdelta=0.01, packet_len=4000)
        for c in range(nch):
            ref = orcs[c].demod(iqs[c, a:b], packet_len=4000, xdelta=0.01)
            d = bool(props[c]["differentialDecoding"])
            assert_parity(got[c], ref, differential=d if a == 0 else (False if np.isfinite(ref["soft"]).all() else d), tag=f"sync call {a}:{b} ch{c}")
    if os.environ.get("PSKD_TP") == "1" and os.environ.get("PSKD_FUSED") == "0":
        assert bank.stats()["tp_packets"] > 0             # every slab ran its own time-parallel plan
    # (b) two pipelined calls, one sync at the end
    bank2 = pk.Bank(nch, props)
    cap = n // 8 + 16
    outs = []
    bufs = [np.ascontiguousarray(iqs[:, :n]), np.ascontiguousarray(iqs[:, n:])]
    for j in range(2):
        o = dict(soft=np.zeros((nch, cap), np.complex64), phase=np.zeros((nch, cap), np.float32), sidx=np.zeros((nch, cap), np.int16),
                 bits=np.zeros((nch, 3 * cap), np.int16), hard=np.zeros((nch, cap), np.uint8))
        rc, ns, nb = bank2.process_raw(bufs[j].ctypes.data, n, n, o["soft"].ctypes.data, o["bits"].ctypes.data, o["phase"].ctypes.data,
                                       o["sidx"].ctypes.data, cap, 3 * cap, xdelta=0.01, packet_len=4000,
                                       flags=B.FLAG_HOST_BUFFERS | B.FLAG_NO_SYNC, hard_ptr=o["hard"].ctypes.data)
        assert rc == 0
        o["ns"], o["nb"] = ns, nb
        outs.append(o)
    bank2.sync()
    orcs = [oracle_built.OracleComponent(**p) for p in props]
    for j, o in enumerate(outs):
        for c in range(nch):
            k, b = int(o["ns"][c]), int(o["nb"][c])
            got = dict(soft=o["soft"][c, :k], phase=o["phase"][c, :k], sidx=o["sidx"][c, :k], bits=o["bits"][c, :b], hard=o["hard"][c, :k])
            ref = orcs[c].demod(bufs[j][c], packet_len=4000, xdelta=0.01)
            d = bool(props[c]["differentialDecoding"])
            assert_parity(got, ref, differential=d if j == 0 else (False if np.isfinite(ref["soft"]).all() else d), tag=f"pipelined call {j} ch{c}")


def test_bank_with_one_stalled_channel(oracle_built):
    """a window shrink on ONE channel of a bank: that channel stalls (and recovers) while its neighbours keep emitting"""
    import psk_soft_b200 as pk
    props = [dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50) for _ in range(3)]
    iqs = np.stack([siggen.gen_shaped(90000, 8, 8, seed=500 + c, sigma=0.03, freq=1e-5 * (c - 1), timing_shift=c) for c in range(3)])
    bank = pk.Bank(3, props)
    orcs = [oracle_built.OracleComponent(**p) for p in props]
    script = [(0, 30000, {}), (30000, 30400, dict(numAvg=40)), (30400, 31000, {}), (31000, 60000, dict(numAvg=250)), (60000, 90000, {})]
    for a, b, ch in script:
        if ch:
            bank.set_props(1, **ch)
            orcs[1].configure(**ch)
        got = bank.process_host(iqs[:, a:b].copy(), xdelta=0.01, packet_len=8000)
        for c in range(3):
            ref = orcs[c].demod(iqs[c, a:b], packet_len=8000, xdelta=0.01)
            assert_parity(got[c], ref, tag=f"{a}:{b} ch{c} {ch}")
            if c == 1 and a in (30000, 30400):
                assert len(ref["sidx"]) == 0          # stalled
            else:
                assert len(ref["sidx"]) > 0
