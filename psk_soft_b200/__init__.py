"""psk_soft_b200 -- B200-native PSK soft demodulator (drop-in for the demod core behind
rh.psk_soft's psk_soft_i::serviceFunction, reference: cpp/psk_soft.cpp:346-618).

The package holds only what that one path needs:
  csrc/      hand-written sm_100a CUDA kernels + the C-ABI host runtime (include/pskd.h)
  host/      C++ host-side mirror of the reference component's interface
  binding.py ctypes view of the C ABI
  bank.py    thin Python mirror (channel bank + single-component view) used by tests and bench

There is no CPU implementation in here; importing `Bank`/`PskSoft` and calling them without
the built CUDA library or without a GPU raises.
"""
from .binding import PskdError, load, lib_path  # noqa: F401
from .bank import Bank, PskSoft, synth_fill, default_props  # noqa: F401

__all__ = ["Bank", "PskSoft", "PskdError", "load", "lib_path", "synth_fill", "default_props"]
