"""Re-run one randomized configuration of tests/test_gpu_parity.py::test_fused_randomized_configurations many times and
report where the GPU result differs from the oracle / from the first GPU run.  usage: fuzz_repro.py SEED [REPEATS]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["PSKD_FUSED"] = "1"; os.environ["PSKD_TP"] = "0"
import siggen
import psk_soft_b200 as pk
from oracle import oracle

seed = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rs = np.random.RandomState(1000 + seed)
nch = 5
props, iqs = [], []
n = int(rs.randint(30000, 70000))
for c in range(nch):
    S = int(rs.choice([8, 9, 10, 16])); M = int(rs.choice([2, 4, 8, 8])); D = int(rs.randint(0, 2))
    A = int(rs.choice([1, 2, 7, 31, 32, 33, 100, 129, 130, 200, 256])); P = int(rs.choice([2, 3, 25, 50, 52, 53, 100, 128]))
    props.append(dict(samplesPerBaud=S, constelationSize=M, numAvg=A, phaseAvg=P, differentialDecoding=D))
    amp = float(10.0 ** rs.uniform(-3, 3))
    x = siggen.gen_shaped(n, S, M, seed=int(rs.randint(1 << 30)), sigma=0.02 * amp, freq=float(rs.uniform(-3e-5, 3e-5)),
                          phase0=float(rs.uniform(0, 6.28)), timing_shift=int(rs.randint(0, S)), amp=amp)
    if rs.rand() < 0.4:
        a0 = int(rs.randint(0, n - 3000)); ln = int(rs.randint(10, 3000)); x[a0:a0 + ln] = 0
        print(f"ch{c}: silent stretch samples {a0}..{a0 + ln}")
    iqs.append(x)
iqs = np.stack(iqs)
pkt = int(rs.choice([97, 640, 1000, 4096, 16000, 64000]))
cuts = sorted(set([0, n] + [int(v) for v in rs.randint(1, n, size=int(rs.randint(0, 3)))]))
print("n", n, "pkt", pkt, "cuts", cuts)
for c in range(nch): print(c, props[c])
refs = []
orcs = [oracle.OracleComponent(**p) for p in props]
for a, b in zip(cuts[:-1], cuts[1:]):
    refs.append([orcs[c].demod(iqs[c, a:b], packet_len=pkt, xdelta=0.01) for c in range(nch)])
first = None
for r in range(reps):
    bank = pk.Bank(nch, props)
    outs = [bank.process_host(iqs[:, a:b].copy(), xdelta=0.01, packet_len=pkt) for a, b in zip(cuts[:-1], cuts[1:])]
    st = bank.stats()
    for ci, out in enumerate(outs):
        for c in range(nch):
            g, f = out[c], refs[ci][c]
            for k in ("sidx", "bits", "phase", "soft"):
                ga, fa = g[k], f[k]
                if k in ("phase", "soft"):
                    with np.errstate(invalid="ignore"):
                        bad = ~(np.abs(ga - fa) <= 1e-4 * np.maximum(1, np.abs(fa))) & np.isfinite(fa)
                else:
                    bad = ga != fa
                if bad.any():
                    idx = np.nonzero(bad)[0]
                    print(f"run {r} call {ci} ch{c} {k}: {len(idx)} differ, first {idx[:6]}  got {ga[idx[:3]]} ref {fa[idx[:3]]}  stats {st}")
                    if k == "bits":
                        bpb = max(1, len(fa) // max(1, len(f["sidx"])))
                        for sym in sorted(set(int(i) // bpb for i in idx[:12]))[:4]:
                            print(f"   sym {sym}: soft got {g['soft'][sym]!r} ref {f['soft'][sym]!r} phase got {g['phase'][sym]!r} ref {f['phase'][sym]!r} sidx {g['sidx'][sym]} bits got {ga[sym*bpb:(sym+1)*bpb]} ref {fa[sym*bpb:(sym+1)*bpb]}")
                    if k == "phase":
                        i0 = max(0, idx[0] - 4)
                        print("   got", ga[i0:i0 + 10]); print("   ref", fa[i0:i0 + 10])
print("done")
