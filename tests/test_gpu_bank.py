"""Channel-bank tests with DEVICE-resident buffers (the `value` path of bench.py): a synthetic
bank generated in HBM, processed through pskd_process with device pointers; sampled channels
are checked against the oracle, the rest through size-independent properties."""
import numpy as np
import pytest

from parity import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fused", "staged", "tp", "legacy"], autouse=True)
def fused_mode(request, monkeypatch):
    """every test of this module runs through the fused kernel, the staged kernels (k_fzs_front +
    k_fzs_cb), the staged kernels with the time-parallel chain, and the legacy staged kernels
    (PSKD_FZS=0: k_front_t + k_chain_par + k_back_par, time-parallel chain on)"""
    monkeypatch.setenv("PSKD_FUSED", "1" if request.param == "fused" else "0")
    monkeypatch.setenv("PSKD_TP", "1" if request.param in ("tp", "legacy") else "0")
    monkeypatch.setenv("PSKD_FZS", "0" if request.param == "legacy" else "1")
    return request.param


def _run_bank(pk, torch, props, nch, n, cuts, seed, packet_len=16000):
    from psk_soft_b200 import binding as B
    S = props["samplesPerBaud"]
    cap = n // S + 8
    iq = torch.empty((nch, n, 2), dtype=torch.float32, device="cuda")
    pk.synth_fill(iq.data_ptr(), n, 0, nch, n, seed=seed, samplesPerBaud=S, constelationSize=props["constelationSize"],
                  sigma=0.02, freq_max=2e-5, pn_sigma=0.01)
    torch.cuda.synchronize()
    bank = pk.Bank(nch, props)
    outs = dict(soft=torch.zeros((nch, cap, 2), dtype=torch.float32, device="cuda"),
                phase=torch.zeros((nch, cap), dtype=torch.float32, device="cuda"),
                sidx=torch.zeros((nch, cap), dtype=torch.int16, device="cuda"),
                bits=torch.zeros((nch, cap * 3), dtype=torch.int16, device="cuda"))
    bpb = {2: 1, 4: 2, 8: 3}[props["constelationSize"]]
    done = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        rc, ns, nb = bank.process_raw(iq.data_ptr() + a * 8, n, b - a,
                                      outs["soft"].data_ptr() + done * 8, outs["bits"].data_ptr() + done * bpb * 2,
                                      outs["phase"].data_ptr() + done * 4, outs["sidx"].data_ptr() + done * 2,
                                      cap, cap * 3, xdelta=0.01, packet_len=packet_len)
        assert rc == 0 and (ns == ns[0]).all()
        done += int(ns[0])
    torch.cuda.synchronize()
    return iq, outs, done, bank


def test_bank_device_buffers_vs_oracle_and_streaming(oracle_built):
    import torch
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50, differentialDecoding=0)
    nch, n = 512, 160000
    iq, one, K, bank = _run_bank(pk, torch, props, nch, n, [0, n], seed=77)
    assert K == n // 8 - 99                                   # floor(N/S) - numAvg + 1 (cpp/psk_soft.cpp:454-457)
    st = bank.stats()
    assert st["symbols_out"] == nch * K and st["seq_channels"] == 0
    # same bank fed in three calls cut at packet boundaries: identical packets -> identical results
    _, three, K3, _ = _run_bank(pk, torch, props, nch, n, [0, 48000, 112000, n], seed=77)
    assert K3 == K
    for k in ("sidx", "bits"):
        assert torch.equal(one[k], three[k]), k
    for k in ("phase", "soft"):
        assert torch.equal(one[k][:, :K], three[k][:, :K]), k
    # sampled channels against the oracle
    iq_h = iq.cpu().numpy().view(np.complex64).reshape(nch, n)
    for c in (0, 1, 255, 511):
        ref = oracle_built.OracleComponent(**props).demod(iq_h[c], packet_len=16000, xdelta=0.01)
        got = dict(soft=one["soft"][c, :K].cpu().numpy().view(np.complex64).reshape(-1), phase=one["phase"][c, :K].cpu().numpy(),
                   sidx=one["sidx"][c, :K].cpu().numpy(), bits=one["bits"][c, :3 * K].cpu().numpy())
        assert_parity(got, ref, tag=f"channel {c}")


def test_bank_differential_qpsk_vs_oracle(oracle_built):
    import torch
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=10, constelationSize=4, numAvg=50, phaseAvg=25, differentialDecoding=1)
    nch, n = 96, 100000
    iq, out, K, _ = _run_bank(pk, torch, props, nch, n, [0, n], seed=5, packet_len=6400)
    iq_h = iq.cpu().numpy().view(np.complex64).reshape(nch, n)
    for c in (0, 37, 95):
        ref = oracle_built.OracleComponent(**props).demod(iq_h[c], packet_len=6400, xdelta=0.01)
        got = dict(soft=out["soft"][c, :K].cpu().numpy().view(np.complex64).reshape(-1), phase=out["phase"][c, :K].cpu().numpy(),
                   sidx=out["sidx"][c, :K].cpu().numpy(), bits=out["bits"][c, :2 * K].cpu().numpy())
        assert_parity(got, ref, differential=True, tag=f"channel {c}")


@pytest.mark.parametrize("S,M,D", [(8, 8, 0), (9, 4, 1), (10, 2, 0)])
def test_bank_unaligned_rows_vs_oracle(oracle_built, S, M, D):
    """odd row lengths and an odd first sample: the input rows of odd channels and every output row start
    8-byte (not 16-byte) aligned -> the 8-byte copy path of the fused kernel's staged blocks and the
    element-wise output stores instead of the 128-bit ones"""
    import torch
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=S, constelationSize=M, numAvg=60, phaseAvg=30, differentialDecoding=D)
    nch, n = 48, 90001
    cap = n // S + 9                                          # odd for every S used here -> odd output rows
    cap += (cap + 1) % 2
    iq = torch.empty((nch, n, 2), dtype=torch.float32, device="cuda")
    pk.synth_fill(iq.data_ptr(), n, 0, nch, n, seed=31 + S, samplesPerBaud=S, constelationSize=M,
                  sigma=0.02, freq_max=2e-5, pn_sigma=0.01)
    torch.cuda.synchronize()
    bank = pk.Bank(nch, props)
    bpb = {2: 1, 4: 2, 8: 3}[M]
    soft = torch.zeros((nch, cap, 2), dtype=torch.float32, device="cuda")
    phase = torch.zeros((nch, cap), dtype=torch.float32, device="cuda")
    sidx = torch.zeros((nch, cap), dtype=torch.int16, device="cuda")
    bits = torch.zeros((nch, cap * 3), dtype=torch.int16, device="cuda")
    skip = 3                                                  # start at an odd sample of every row
    rc, ns, nb = bank.process_raw(iq.data_ptr() + skip * 8, n, n - skip, soft.data_ptr(), bits.data_ptr(), phase.data_ptr(),
                                  sidx.data_ptr(), cap, cap * 3, xdelta=0.01, packet_len=7000)
    assert rc == 0 and (ns == ns[0]).all()
    K = int(ns[0])
    torch.cuda.synchronize()
    iq_h = iq.cpu().numpy().view(np.complex64).reshape(nch, n)
    for c in (0, 1, 22, 47):
        ref = oracle_built.OracleComponent(**props).demod(iq_h[c, skip:], packet_len=7000, xdelta=0.01)
        got = dict(soft=soft[c, :K].cpu().numpy().view(np.complex64).reshape(-1), phase=phase[c, :K].cpu().numpy(),
                   sidx=sidx[c, :K].cpu().numpy(), bits=bits[c, :bpb * K].cpu().numpy())
        assert_parity(got, ref, differential=bool(D), tag=f"S={S} channel {c}")


def test_full_size_bank_properties(oracle_built, fused_mode):
    """BASELINE.json configs[3] at FULL size (QPSK, S=8, 4096 channels x 1M samples, packets 64000) through
    size-independent properties: symbol counts, value ranges, one call == two calls cut at a packet boundary,
    a 64-channel sub-bank processed alone (other launch geometry / kernel choice) == the same channels inside the
    bank, and sampled channels against the oracle over their full length."""
    import torch
    import psk_soft_b200 as pk
    if torch.cuda.mem_get_info()[0] < 70e9:
        pytest.skip("needs ~60 GB of free device memory")
    props = dict(samplesPerBaud=8, constelationSize=4, numAvg=100, phaseAvg=50, differentialDecoding=0)
    nch, n, pkt = 4096, 1_000_000, 64000
    iq, one, K, bank = _run_bank(pk, torch, props, nch, n, [0, n], seed=4, packet_len=pkt)
    assert K == n // 8 - 99
    st = bank.stats()
    assert st["symbols_out"] == nch * K
    sidx = one["sidx"][:, :K]
    assert int(sidx.min()) >= 0 and int(sidx.max()) < 8
    bits = one["bits"][:, :2 * K]
    assert int(bits.min()) >= 0 and int(bits.max()) <= 1
    assert bool(torch.isfinite(one["phase"][:, :K]).all()) and bool(torch.isfinite(one["soft"][:, :K]).all())
    del bank
    # two calls cut at a packet boundary: identical packets -> identical integers, floats within the parity tolerance
    _, two, K2, _ = _run_bank(pk, torch, props, nch, n, [0, 7 * pkt, n], seed=4, packet_len=pkt)
    assert K2 == K
    assert torch.equal(one["sidx"], two["sidx"]) and torch.equal(one["bits"], two["bits"])
    for k in ("phase", "soft"):
        a, b = one[k][:, :K].float(), two[k][:, :K].float()
        assert bool(((a - b).abs() <= 1e-4 * torch.clamp(a.abs(), min=1.0)).all()), k
    del two
    # a sub-bank processed alone
    lo, hi = 1000, 1064
    sub = pk.Bank(hi - lo, props)
    cap = n // 8 + 8
    s_soft = torch.zeros((hi - lo, cap, 2), dtype=torch.float32, device="cuda")
    s_phase = torch.zeros((hi - lo, cap), dtype=torch.float32, device="cuda")
    s_sidx = torch.zeros((hi - lo, cap), dtype=torch.int16, device="cuda")
    s_bits = torch.zeros((hi - lo, cap * 3), dtype=torch.int16, device="cuda")
    rc, ns, nb = sub.process_raw(iq[lo:hi].data_ptr(), n, n, s_soft.data_ptr(), s_bits.data_ptr(), s_phase.data_ptr(),
                                 s_sidx.data_ptr(), cap, cap * 3, xdelta=0.01, packet_len=pkt)
    torch.cuda.synchronize()
    assert rc == 0 and int(ns[0]) == K
    assert torch.equal(s_sidx[:, :K], one["sidx"][lo:hi, :K]) and torch.equal(s_bits[:, :2 * K], one["bits"][lo:hi, :2 * K])
    a, b = s_phase[:, :K], one["phase"][lo:hi, :K]
    assert bool(((a - b).abs() <= 1e-4 * torch.clamp(a.abs(), min=1.0)).all())
    # sampled channels against the oracle, full length
    for c in (0, 1777, 4095):
        iq_h = iq[c].cpu().numpy().view(np.complex64).reshape(n)
        ref = oracle_built.OracleComponent(**props).demod(iq_h, packet_len=pkt, xdelta=0.01)
        got = dict(soft=one["soft"][c, :K].cpu().numpy().view(np.complex64).reshape(-1), phase=one["phase"][c, :K].cpu().numpy(),
                   sidx=one["sidx"][c, :K].cpu().numpy(), bits=one["bits"][c, :2 * K].cpu().numpy())
        assert_parity(got, ref, tag=f"full-size channel {c}")
