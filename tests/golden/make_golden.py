"""Generate the committed golden vectors from the UNMODIFIED reference build (oracle/_ref).

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz: the seeded input stream (complex64), the properties / packetisation and
the four out-port payloads the reference produced.  These pin the C restatement (CPU tests) and
the CUDA path (GPU tests) on machines where the reference sources do not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import siggen  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {
    # name: (generator kwargs, props, packet_len, xdelta)
    "qpsk_s8_coherent": (dict(n=40000, S=8, M=4, seed=1, sigma=0.02, freq=1e-5, shift=3),
                         dict(samplesPerBaud=8, constelationSize=4, numAvg=100, phaseAvg=50, differentialDecoding=0), 16000, 0.01),
    "bpsk_s10_offset_pn": (dict(n=50000, S=10, M=2, seed=2, sigma=0.05, freq=1e-4, pn=0.002, shift=4),
                           dict(samplesPerBaud=10, constelationSize=2, numAvg=100, phaseAvg=50, differentialDecoding=0), 6400, 0.01),
    "psk8_s8_diff": (dict(n=40000, S=8, M=8, seed=3, sigma=0.02, freq=2e-5, shift=1),
                     dict(samplesPerBaud=8, constelationSize=8, numAvg=100, phaseAvg=50, differentialDecoding=1), 8000, 0.01),
    "psk8_s8_coherent_xd1": (dict(n=40000, S=8, M=8, seed=4, sigma=0.02, freq=2e-5, shift=6),
                             dict(samplesPerBaud=8, constelationSize=8, numAvg=37, phaseAvg=20, differentialDecoding=0), 1001, 1.0),
    "bpsk_s9_diff_smallpkt": (dict(n=30000, S=9, M=2, seed=5, sigma=0.05, freq=0.0, shift=2),
                              dict(samplesPerBaud=9, constelationSize=2, numAvg=64, phaseAvg=50, differentialDecoding=1), 777, 0.01),
    "qpsk_lowsnr": (dict(n=60000, S=8, M=4, seed=6, sigma=0.25, freq=5e-5, shift=0),
                    dict(samplesPerBaud=8, constelationSize=4, numAvg=100, phaseAvg=50, differentialDecoding=0), 16000, 0.01),
}


def make_input(g):
    return siggen.gen_shaped(g["n"], g["S"], g["M"], seed=g["seed"], sigma=g["sigma"], freq=g["freq"],
                             pn_sigma=g.get("pn", 0.0), timing_shift=g["shift"])


def main():
    oracle.build(ref=True)
    assert oracle.have_ref(), "needs /root/reference to build oracle/_ref"
    for name, (g, props, pkt, xd) in CASES.items():
        iq = make_input(g)
        out = oracle.RefComponent(**props).demod(iq, packet_len=pkt, xdelta=xd)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), iq=iq, packet_len=pkt, xdelta=xd,
                            props=np.array([props[k] for k in oracle.PROPS[:5]], dtype=np.int64),
                            soft=out["soft"], bits=out["bits"], phase=out["phase"], sidx=out["sidx"])
        print(name, len(out["phase"]), "symbols")
    # the six cases of the reference's own test module (tests/test_psk_soft.py:160-176), seed 100
    for cname, c in siggen.reference_cases().items():
        props = dict(samplesPerBaud=8, constelationSize=c["M"], numAvg=100, phaseAvg=50, differentialDecoding=int(c["differential"]))
        out = oracle.RefComponent(**props).demod(c["iq"], packet_len=64000, xdelta=0.01)
        np.savez_compressed(os.path.join(HERE, "reftest_" + cname + ".npz"), iq=c["iq"], packet_len=64000, xdelta=0.01,
                            props=np.array([props[k] for k in oracle.PROPS[:5]], dtype=np.int64),
                            soft=out["soft"], bits=out["bits"], phase=out["phase"], sidx=out["sidx"], syms=c["syms"])
        print("reftest", cname, len(out["phase"]), "symbols")


if __name__ == "__main__":
    main()
