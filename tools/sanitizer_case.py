import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import siggen
import psk_soft_b200 as pk
mode = sys.argv[1]
os.environ["PSKD_FUSED"] = "1" if mode == "fused" else "0"
os.environ["PSKD_TP"] = "1" if mode == "tp" else "0"
rs = np.random.RandomState(3)
nch, nmax = 6, 30000
props, lens, iqs = [], [], np.zeros((nch, nmax), np.complex64)
for c in range(nch):
    S = int(rs.choice([8, 9, 10])); M = int(rs.choice([2, 4, 8])); A = int(rs.choice([50, 100])); P = int(rs.choice([25, 50])); D = int(rs.randint(0, 2))
    n = int(rs.randint(nmax // 2, nmax + 1))
    props.append(dict(samplesPerBaud=S, constelationSize=M, numAvg=A, phaseAvg=P, differentialDecoding=D)); lens.append(n)
    iqs[c, :n] = siggen.gen_shaped(n, S, M, seed=100 + c, sigma=0.05, freq=3e-5, timing_shift=c)
bank = pk.Bank(nch, props)
out = bank.process_host(iqs, n_complex=lens, xdelta=0.01, packet_len=2000)
out2 = bank.process_host(iqs[:, :5000].copy(), n_complex=[5000] * nch, xdelta=0.01, packet_len=700)
print(mode, "ok", sum(len(o["sidx"]) for o in out), bank.stats())
