"""Several GPUs driven from ONE process (include/pskd.h: one handle per GPU, handles may be driven from different host
threads): per-device kernel configuration, and the C++ host mirror's box runner (psk_box_gpu) sharding one bank over
every visible GPU.  The multi-device cases skip on a single-GPU box; the box demo runs on however many GPUs there are."""
import os
import subprocess

import numpy as np
import pytest

import siggen
from parity import assert_parity

pytestmark = pytest.mark.gpu


def test_box_demo_shards_one_bank_over_all_gpus():
    """psk_box_gpu (psk_soft_b200/host/psk_soft_gpu.hpp) over all visible GPUs: per-GPU symbol counts add up and
    sampled channels of every shard equal single-channel runs"""
    from psk_soft_b200 import _build
    exe = os.path.join(_build.LIB_DIR, "demo_box")
    if not os.path.isfile(exe):
        _build.build_host_demo()
    res = subprocess.run([exe, "768", "40000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "box ok" in res.stdout, res.stdout


@pytest.mark.parametrize("mode", ["legacy_staged", "fused", "staged"])
def test_two_devices_in_one_process(mode, oracle_built, monkeypatch):
    """banks on two devices in one process, kernels that need the > 48 KB dynamic shared memory opt-in
    (k_front_t<10> ~60 KB, k_fused<16> ~66 KB): function attributes are per device, so every device a process
    launches on must be configured (it used to be only the first)"""
    import torch
    import psk_soft_b200 as pk
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("PSKD_FUSED", "1" if mode == "fused" else "0")
    monkeypatch.setenv("PSKD_FZS", "0" if mode == "legacy_staged" else "1")
    props = [dict(samplesPerBaud=10, constelationSize=8, numAvg=100, phaseAvg=50),
             dict(samplesPerBaud=16, constelationSize=4, numAvg=64, phaseAvg=100, differentialDecoding=1),
             dict(samplesPerBaud=9, constelationSize=2, numAvg=50, phaseAvg=25)]
    n = 60000
    iq = np.stack([siggen.gen_shaped(n, p["samplesPerBaud"], p["constelationSize"], seed=70 + c, sigma=0.03, freq=2e-5, timing_shift=c)
                   for c, p in enumerate(props)])
    ndev = min(torch.cuda.device_count(), 4)
    banks = [pk.Bank(3, props, device=d) for d in range(ndev)]
    for d in reversed(range(ndev)):                      # the LAST device first: it must not depend on device 0's set-up
        got = banks[d].process_host(iq, xdelta=0.01, packet_len=8000)
        for c, p in enumerate(props):
            ref = oracle_built.OracleComponent(**p).demod(iq[c], packet_len=8000, xdelta=0.01)
            assert_parity(got[c], ref, differential=bool(p.get("differentialDecoding", 0)), tag=f"device {d} channel {c} ({mode})")
