#!/usr/bin/env python
"""step_trace.py -- per-step times of the bench workload over a long run, with NVML clocks / power sampled every few ms.

    python tools/probe/step_trace.py [--channels 4096] [--steps 60] [--rounds 3]

Question it answers: does the step time of the fused kernel drift with time under load (power management), and does the
profiled pass of bench.py (CUDA events around every launch) see another kernel time than the plain one?  Round r runs
`steps` steps with per-launch profiling off (even r) or on (odd r); a CUDA event after every step gives the step times.
Prints one JSON line per round.
"""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


class Nvml(threading.Thread):
    def __init__(self, dev, period=0.004):
        super().__init__(daemon=True)
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        self.period = period
        self.rows = []
        self.go = True

    def run(self):
        nv = self.nv
        while self.go:
            try:
                self.rows.append((time.time(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                                  else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception as e:          # noqa: BLE001
                self.rows.append((time.time(), -1, -1.0, -1))
            time.sleep(self.period)

    def energy_j(self):
        try:
            return self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h) / 1e3
        except Exception:                    # noqa: BLE001
            return None

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 <= x[0] <= t1 and x[1] > 0]
        if not r:
            return None
        sm = sorted(x[1] for x in r)
        reasons = 0
        for x in r:
            reasons |= x[3]
        return {"n": len(r), "sm_min": sm[0], "sm_med": sm[len(sm) // 2], "sm_max": sm[-1],
                "w_max": max(x[2] for x in r), "w_med": sorted(x[2] for x in r)[len(r) // 2], "reasons_or": hex(reasons)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--pause", type=float, default=0.0, help="idle seconds between rounds")
    ap.add_argument("--no-profiled", action="store_true", help="every round with per-launch profiling off")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    import psk_soft_b200 as pk
    from psk_soft_b200 import binding as B
    nch, n = a.channels, a.samples
    props = dict(samplesPerBaud=8, numAvg=100, constelationSize=8, phaseAvg=50, differentialDecoding=0)
    cap = n // 8 + 8
    iq = torch.empty((nch, n, 2), dtype=torch.float32, device="cuda")
    soft = torch.empty((nch, cap, 2), dtype=torch.float32, device="cuda")
    phase = torch.empty((nch, cap), dtype=torch.float32, device="cuda")
    sidx = torch.empty((nch, cap), dtype=torch.int16, device="cuda")
    bits = torch.empty((nch, cap * 3), dtype=torch.int16, device="cuda")
    pk.synth_fill(iq.data_ptr(), n, 0, nch, n, seed=4, samplesPerBaud=8, constelationSize=8, sigma=0.02, freq_max=2e-5,
                  pn_sigma=0.0, device=0, period=n)
    torch.cuda.synchronize()
    bank = pk.Bank(nch, [props] * nch, device=0)
    stream = torch.cuda.ExternalStream(bank.stream, device=0)

    def step():
        bank.process_raw(iq.data_ptr(), n, n, soft.data_ptr(), bits.data_ptr(), phase.data_ptr(), sidx.data_ptr(), cap, cap * 3,
                         xdelta=0.01, packet_len=64000, flags=B.FLAG_NO_SYNC, counts=False)

    mon = Nvml(0)
    mon.start()
    for _ in range(3):
        step()
    bank.sync()
    for r in range(a.rounds):
        prof = bool(r & 1) and not a.no_profiled
        bank.profile_read(reset=True)
        bank.profile_enable(prof)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        t0 = time.time()
        e0 = mon.energy_j()
        evs[0].record(stream)
        for i in range(a.steps):
            step()
            evs[i + 1].record(stream)
        bank.sync()
        torch.cuda.synchronize()
        t1 = time.time()
        e1 = mon.energy_j()
        ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
        kern = bank.profile_read(reset=True) if prof else {}
        out = {"tag": a.tag, "round": r, "profiled": prof, "channels": nch, "steps": a.steps,
               "ms_first5": [round(x, 3) for x in ms[:5]], "ms_last5": [round(x, 3) for x in ms[-5:]],
               "ms_mean_2nd_half": round(sum(ms[len(ms) // 2:]) / (len(ms) - len(ms) // 2), 3),
               "joule_per_step": (round((e1 - e0) / a.steps, 3) if e0 is not None and e1 is not None else None),
               "watt_avg": (round((e1 - e0) / (t1 - t0), 1) if e0 is not None and e1 is not None else None),
               "ms_min": round(min(ms), 3), "ms_med": round(sorted(ms)[len(ms) // 2], 3), "ms_max": round(max(ms), 3),
               "ms_every10": [round(x, 2) for x in ms[::10]],
               "kernel_ms_per_launch": {k: round(v[0] / max(v[1], 1), 4) for k, v in kern.items()},
               "nvml": mon.window(t0, t1)}
        print(json.dumps(out), flush=True)
        if a.pause > 0:
            time.sleep(a.pause)
    mon.go = False


if __name__ == "__main__":
    main()
