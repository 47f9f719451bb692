"""No-GPU checks of the drop-in boundary: the C-ABI library builds/loads here, exports every symbol
include/pskd.h declares, keeps the reference's property defaults, and FAILS LOUDLY (no CPU
fallback) when asked to compute without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pskd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pskd_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import psk_soft_b200 as pk
    from psk_soft_b200 import binding
    lib = pk.load()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pskd.h but not exported by libpskd.so"
    assert set(binding.EXPORTS) == set(names)
    assert lib.pskd_abi_version() == 2


def test_default_properties_match_reference_prf():
    """psk_soft.prf.xml:23-60 / cpp/psk_soft_base.cpp:94-150"""
    import psk_soft_b200 as pk
    assert pk.default_props() == dict(samplesPerBaud=10, numAvg=100, constelationSize=4, phaseAvg=50,
                                      differentialDecoding=0, resetState=0)


def test_struct_layouts_match_header():
    from psk_soft_b200 import binding as B
    assert ctypes.sizeof(B.Props) == 16       # uint16, uint32, uint16, uint16, uint8, uint8 with natural alignment
    assert B.Props.numAvg.offset == 4 and B.Props.differentialDecoding.offset == 12
    assert ctypes.sizeof(B.Input) == 64 and ctypes.sizeof(B.Output) == 72
    assert ctypes.sizeof(B.KernelTime) == 56 and ctypes.sizeof(B.Synth) == 32


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import psk_soft_b200 as pk
    with pytest.raises(pk.PskdError) as e:
        pk.Bank(2)
    assert e.value.code == -2          # PSKD_ERR_CUDA
    with pytest.raises(pk.PskdError):
        pk.PskSoft(samplesPerBaud=8).push(np.ones(1000, np.complex64))


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under psk_soft_b200/ may import, load or link it"""
    banned = ("import oracle", "from oracle", "libpsk_oracle", "libpsk_ref", "psk_oracle", "ref_driver", "orc_", "ref_demod")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "psk_soft_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for b in banned:
                    assert b not in txt, f"{os.path.join(dirpath, f)} references the oracle ({b})"


def test_bench_clock_sampler_degrades_without_a_gpu():
    """bench.py's clock / energy sampler (NVML every 5 ms, nvidia-smi as the fallback) must not take the bench down on a box
    that has neither: it reports that it has no samples"""
    import importlib.util
    import time
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.1)
    s.stop()
    out = s.summary(time.time() - 1.0, time.time())
    assert "reasons" in out and "sm_mhz" in out
    if out["sm_mhz"] is None:
        assert s.energy_j() is None or isinstance(s.energy_j(), float)
