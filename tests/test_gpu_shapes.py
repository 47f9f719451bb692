"""Parity at the shapes that are BENCHMARKED (bench.py workloads): the device-resident bank goes through
pskd_process exactly as in bench.py's `value` leg (automatic kernel choice), sampled channels are compared with the
oracle over their full length, the rest through size-independent properties.  Where oracle/_ref (the unmodified
reference build) is present, the sampled channels are checked against IT as well as against the C port."""
import os

import numpy as np
import pytest

from parity import assert_parity

pytestmark = pytest.mark.gpu
BPB = {2: 1, 4: 2, 8: 3}


def _checkers(oracle):
    out = [("port", oracle.OracleComponent)]
    if oracle.have_ref():
        out.append(("reference", oracle.RefComponent))
    return out


def _bank_run(pk, torch, table, n, seed, packet_len, pn=0.0, sigma=0.02, freq_max=2e-5, calls=1, no_hard_calls=()):
    """generate the bank class by class in HBM (as bench.py does), run `calls` pskd_process calls over it.  Calls listed in
    `no_hard_calls` do not ask for the additional packed hard-symbol output -- exactly bench.py's call: a uniform bank then
    runs the k_fused instance with its back-stage variant inlined (k_fused<S,PC,CT,BK>), the others the run-time dispatch."""
    nch = len(table)
    Smin = min(p["samplesPerBaud"] for p in table)
    cap = n // Smin + 8
    torch.cuda.empty_cache()
    iq = torch.empty((nch, n, 2), dtype=torch.float32, device="cuda")
    c0 = 0
    while c0 < nch:
        key = (table[c0]["samplesPerBaud"], table[c0]["constelationSize"])
        c1 = c0
        while c1 < nch and (table[c1]["samplesPerBaud"], table[c1]["constelationSize"]) == key:
            c1 += 1
        pk.synth_fill(iq[c0].data_ptr(), n, c0, c1 - c0, n, seed=seed, samplesPerBaud=key[0], constelationSize=key[1],
                      sigma=sigma, freq_max=freq_max, pn_sigma=pn, period=n)
        c0 = c1
    torch.cuda.synchronize()
    bank = pk.Bank(nch, table)
    outs = []
    for j in range(calls):
        o = dict(soft=torch.zeros((nch, cap, 2), dtype=torch.float32, device="cuda"),
                 phase=torch.zeros((nch, cap), dtype=torch.float32, device="cuda"),
                 sidx=torch.zeros((nch, cap), dtype=torch.int16, device="cuda"),
                 bits=torch.zeros((nch, cap * 3), dtype=torch.int16, device="cuda"))
        if j not in no_hard_calls:
            o["hard"] = torch.zeros((nch, cap), dtype=torch.uint8, device="cuda")
        rc, ns, nb = bank.process_raw(iq.data_ptr(), n, n, o["soft"].data_ptr(), o["bits"].data_ptr(), o["phase"].data_ptr(),
                                      o["sidx"].data_ptr(), cap, cap * 3, xdelta=0.01, packet_len=packet_len,
                                      hard_ptr=o["hard"].data_ptr() if "hard" in o else None)
        assert rc == 0
        o["ns"], o["nb"] = ns, nb
        outs.append(o)
    torch.cuda.synchronize()
    return iq, outs, bank


def _channel(o, c):
    k, b = int(o["ns"][c]), int(o["nb"][c])
    d = dict(soft=o["soft"][c, :k].cpu().numpy().view(np.complex64).reshape(-1), phase=o["phase"][c, :k].cpu().numpy(),
             sidx=o["sidx"][c, :k].cpu().numpy(), bits=o["bits"][c, :b].cpu().numpy())
    if "hard" in o:
        d["hard"] = o["hard"][c, :k].cpu().numpy()
    return d


def _check_sampled(oracle, iq, outs, table, rows, packet_len, tag):
    """the sampled channels against the CPU checkers, call after call (the checker carries its state like the bank)"""
    n = iq.shape[1]
    for c in rows:
        iq_h = iq[c].cpu().numpy().view(np.complex64).reshape(n)
        for name, cls in _checkers(oracle):
            comp = cls(**table[c])
            for j, o in enumerate(outs):
                ref = comp.demod(iq_h, packet_len=packet_len, xdelta=0.01)
                d = bool(table[c]["differentialDecoding"])
                assert_parity(_channel(o, c), ref, differential=d if j == 0 else (False if np.isfinite(ref["soft"]).all() else d),
                              tag=f"{tag} channel {c} call {j} vs {name}")


@pytest.mark.parametrize("M", [8, 4], ids=["bank8psk", "config4"])
def test_bench_bank_full_shape(M, oracle_built):
    """bench.py's default workload (8-PSK) and configs[3] (QPSK) at FULL size: 4096 channels x 1M samples, packets of
    64000, two calls over the replayed buffer (the second starts from carried state, as every timed bench step does)."""
    import torch
    import psk_soft_b200 as pk
    import gc
    gc.collect(); torch.cuda.empty_cache()
    if torch.cuda.mem_get_info()[0] < 70e9:
        pytest.skip("needs ~60 GB of free device memory")
    props = dict(samplesPerBaud=8, constelationSize=M, numAvg=100, phaseAvg=50, differentialDecoding=0)
    nch, n, pkt = 4096, 1_000_000, 64000
    table = [props] * nch
    # call 0 asks for the packed hard symbols too (run-time dispatch of the back stage), call 1 is bench.py's call (no hard
    # output: the kernel instance with the bank's back-stage variant inlined, k_fused<8,52,6,BK> -- the benchmarked kernel)
    iq, outs, bank = _bank_run(pk, torch, table, n, seed=4, packet_len=pkt, calls=2, no_hard_calls=(1,))
    K0, K1 = n // 8 - 99, n // 8
    assert (outs[0]["ns"] == K0).all() and (outs[1]["ns"] == K1).all()
    st = bank.stats()
    assert st["symbols_out"] == nch * (K0 + K1) and st["seq_channels"] == 0
    for o, K in zip(outs, (K0, K1)):
        assert int(o["sidx"][:, :K].min()) >= 0 and int(o["sidx"][:, :K].max()) < 8
        b = o["bits"][:, :BPB[M] * K]
        assert int(b.min()) >= 0 and int(b.max()) <= 1
        if M == 8:
            assert 0.45 < float(b.float().mean()) < 0.55        # real decisions, not a constant
        assert bool(torch.isfinite(o["phase"][:, :K]).all()) and bool(torch.isfinite(o["soft"][:, :K]).all())
    # 16 sampled channels over their full length, both calls: first / last channel, neighbours, spread
    rows = sorted(set([0, 1, 2, 2047, 2048, 4094, 4095] + [int(i * nch / 9) + 3 for i in range(9)]))
    _check_sampled(oracle_built, iq, outs, table, rows, pkt, f"full-size M={M}")


def test_config3_full_shape(oracle_built):
    """configs[2]: 8-PSK, differential, 256 channels x 4M samples (the time-parallel staged kernels)"""
    import torch
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50, differentialDecoding=1)
    nch, n, pkt = 256, 4_000_000, 64000
    table = [props] * nch
    iq, outs, bank = _bank_run(pk, torch, table, n, seed=3, packet_len=pkt, calls=2)
    assert (outs[0]["ns"] == n // 8 - 99).all() and (outs[1]["ns"] == n // 8).all()
    st = bank.stats()
    assert st["tp_packets"] > 0, st                       # the packets really ran time-parallel
    _check_sampled(oracle_built, iq, outs, table, [0, 1, 100, 255], pkt, "config3")


def test_config2_full_shape(oracle_built):
    """configs[1]: BPSK, S=10, ONE channel x 64M samples, carrier offset (the packet-end wrap fires) + phase wander, all
    1000 packets against the checkers"""
    import torch
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=10, constelationSize=2, numAvg=100, phaseAvg=50, differentialDecoding=0)
    n, pkt = 64_000_000, 64000
    iq, outs, bank = _bank_run(pk, torch, [props], n, seed=2, packet_len=pkt, pn=0.02, sigma=0.05, freq_max=1e-4, calls=1)
    st = bank.stats()
    assert int(outs[0]["ns"][0]) == n // 10 - 99
    assert st["tp_packets"] >= 990 and st["wraps"] > 100, st
    _check_sampled(oracle_built, iq, outs, [props], [0], pkt, "config2")


def _mixed_table(nch, seed, extra=()):
    rs = np.random.RandomState(seed)
    rows = [(int(rs.choice([8, 9, 10])), int(rs.choice([2, 4, 8])), int(rs.choice([50, 100, 200])),
             int(rs.choice([25, 50, 100])), int(rs.randint(0, 2))) for _ in range(nch)]
    rows += list(extra)
    rows.sort(key=lambda r: (r[0], r[1]))
    return [dict(samplesPerBaud=S, constelationSize=M, numAvg=A, phaseAvg=P, differentialDecoding=D) for (S, M, A, P, D) in rows]


@pytest.mark.parametrize("mode", ["auto", "staged"])
def test_config5_mixed_bank(mode, oracle_built, monkeypatch):
    """configs[4] shape at reduced size: a 3000+-channel MIXED bank (S in {8,9,10}, M in {2,4,8}, numAvg in {50,100,200},
    phaseAvg in {25,50,100}, differential on/off) plus a remainder the fused kernel does not take (samplesPerBaud 12,
    phaseAvg 200).  auto: one fused launch per samples-per-symbol class + the staged kernels for the remainder; staged:
    everything through the staged kernels.  16 sampled channels against the checkers."""
    import torch
    import psk_soft_b200 as pk
    monkeypatch.setenv("PSKD_FUSED_MIN", "256")
    if mode == "staged":
        monkeypatch.setenv("PSKD_FUSED", "0")
    extra = [(12, 8, 100, 50, 0)] * 3 + [(8, 4, 100, 200, 0)] * 3 + [(16, 2, 64, 30, 1)] * 2
    table = _mixed_table(3000, 5, extra)
    nch, n, pkt = len(table), 96000, 16000
    iq, outs, bank = _bank_run(pk, torch, table, n, seed=55, packet_len=pkt, pn=0.005, calls=2)
    for j, o in enumerate(outs):
        for c in range(0, nch, 97):
            p = table[c]
            K = (n // p["samplesPerBaud"] - p["numAvg"] + 1) if j == 0 else None
            if K is not None:
                assert int(o["ns"][c]) == K, (c, p)
            assert int(o["nb"][c]) == int(o["ns"][c]) * BPB[p["constelationSize"]]
    kinds = {}
    for c, p in enumerate(table):
        kinds.setdefault((p["samplesPerBaud"], p["phaseAvg"] > 128), []).append(c)
    rows = sorted(set([0, nch - 1] + [v[len(v) // 2] for v in kinds.values()] + [v[0] for v in kinds.values()] +
                      [int(i * nch / 5) + 1 for i in range(5)]))[:20]
    _check_sampled(oracle_built, iq, outs, table, rows, pkt, f"mixed bank ({mode})")


def test_strong_scaled_shard_shape(oracle_built):
    """the per-GPU shard of the 8-GPU strong-scaling run: 512 channels x 1M samples, 8-PSK -- too few channels for the
    fused kernel, so the time-parallel staged kernels (k_fzs_front + k_fzs_cb over packets) carry it"""
    import torch
    import psk_soft_b200 as pk
    props = dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50, differentialDecoding=0)
    nch, n, pkt = 512, 1_000_000, 64000
    table = [props] * nch
    iq, outs, bank = _bank_run(pk, torch, table, n, seed=4, packet_len=pkt, calls=3)
    st = bank.stats()
    assert st["tp_packets"] > 0 and st["seq_channels"] == 0, st      # time-parallel, and no channel fell back to the sequential chain
    _check_sampled(oracle_built, iq, outs, table, [0, 1, 255, 256, 511], pkt, "512-channel shard")
