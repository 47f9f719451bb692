#!/usr/bin/env python
"""bench.py -- headline benchmark of the PSK soft-demod hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ...]

A "step" is one pass of the hot path (pskd_process) over one batch of synthetic input: a channel
bank of `channels` independent IQ channels x `samples` complex samples each, resident in HBM.
Default workload (BASELINE.json north_star target, SURVEY.md 8d config 4 shape): 8-PSK, 8
samples/symbol, numAvg 100, phaseAvg 50, coherent, 4096 channels x 1e6 samples PER GPU, emulated
BULKIO packets of 64000 samples, SRI.xdelta 0.01.  Channels are independent, so N GPUs each run
their own bank with no collective ("scaling": "weak").

Prints ONE JSON line (rank 0).  `value` = Msamples/s with inputs resident in HBM; `e2e` = same
metric through the C ABI with HOST (pinned) buffers, H2D/D2H inside the timed region;
`roofline` = algorithmic bytes / measured kernel time of the dominant kernel vs the measured HBM
copy peak; `cpu_baseline` = the reference's CPU demod (oracle/_ref, else the C port) on a bounded
sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (M, S, A, P, D, channels, samples, sigma, freq_max, pn)
    "bank8psk": dict(M=8, S=8, A=100, P=50, D=0, channels=4096, samples=1_000_000, sigma=0.02, freq_max=2e-5, pn=0.0,
                     desc="8-PSK, 8 samples/symbol, 4096-channel bank x 1M samples per GPU (north_star target; config-4 shape)"),
    "config4": dict(M=4, S=8, A=100, P=50, D=0, channels=4096, samples=1_000_000, sigma=0.02, freq_max=2e-5, pn=0.0,
                    desc="QPSK, 8 samples/symbol, 4096-channel bank x 1M samples per GPU (configs[3])"),
    "config3": dict(M=8, S=8, A=100, P=50, D=1, channels=256, samples=4_000_000, sigma=0.02, freq_max=2e-5, pn=0.0,
                    desc="8-PSK, 8 samples/symbol, differential, 256 channels x 4M samples (configs[2])"),
    "config2": dict(M=2, S=10, A=100, P=50, D=0, channels=1, samples=64_000_000, sigma=0.05, freq_max=1e-4, pn=0.02,
                    desc="BPSK, 10 samples/symbol, single channel, 64M samples, carrier offset + phase noise (configs[1])"),
    "config1": dict(M=4, S=8, A=100, P=50, D=0, channels=1, samples=1_000_000, sigma=0.02, freq_max=1e-5, pn=0.0,
                    desc="QPSK, 8 samples/symbol, single channel, 1M samples (configs[0])"),
}
PACKET_LEN = 64000
XDELTA = 0.01


def bits_per_baud(M):
    return {2: 1, 4: 2, 8: 3}.get(M, 0)


def algorithmic_bytes(w, channels, samples, first_call=False):
    """8*N input + K*(8 soft + 4 phase + 2 sampleIndex + 2*b bits) per channel."""
    K = samples // w["S"] - (w["A"] - 1 if first_call else 0)
    return channels * (8 * samples + K * (8 + 4 + 2 + 2 * bits_per_baud(w["M"])))


def props_of(w):
    return dict(samplesPerBaud=w["S"], numAvg=w["A"], constelationSize=w["M"], phaseAvg=w["P"], differentialDecoding=w["D"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or \
               [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(torch, dev):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function), so that the pinned
    host buffers of the e2e leg are first-touched on the GPU's own NUMA node.  Best effort: returns what it did."""
    try:
        p = torch.cuda.get_device_properties(dev)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        cpulist = open(f"{base}/local_cpulist").read().strip()
        node = open(f"{base}/numa_node").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return {"gpu_pci": bdf, "numa_node": node, "cpus": len(cpus), "bound": True}
        return {"gpu_pci": bdf, "numa_node": node, "cpus": len(allowed), "bound": False}
    except Exception as e:  # no sysfs entry, not permitted, old torch ...
        return {"bound": False, "why": type(e).__name__}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline leg: the reference's own demod core on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_demod_rate(w, iq_host, threads, repeats=1):
    """Times the reference demod (oracle/_ref when present, else the C port) on iq_host
    [channels, samples] complex64 using `threads` host threads (one component per channel).
    Returns (Msamples/s, kind, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    if oracle.have_ref():
        cls, kind = oracle.RefComponent, "reference"
    else:
        if not os.path.isfile(oracle.ORC_SO):
            oracle.build(ref=False)
        cls, kind = oracle.OracleComponent, "port"
    nch, n = iq_host.shape
    comps = [cls(**props_of(w)) for _ in range(nch)]

    def work(c):
        comps[c].demod(iq_host[c], packet_len=PACKET_LEN, xdelta=XDELTA, keep=False)   # ctypes releases the GIL

    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, range(nch)))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return nch * n / best / 1e6, kind, best


def host_sample(w, channels, samples, seed, device_ok):
    """[channels, samples] complex64 of the workload's synthetic input in HOST memory."""
    import numpy as np
    if device_ok:
        import torch
        import psk_soft_b200 as pk
        buf = torch.empty((channels, samples, 2), dtype=torch.float32, device="cuda")
        pk.synth_fill(buf.data_ptr(), samples, 0, channels, samples, seed=seed, samplesPerBaud=w["S"],
                      constelationSize=w["M"], sigma=w["sigma"], freq_max=w["freq_max"], pn_sigma=w["pn"],
                      device=torch.cuda.current_device())
        torch.cuda.synchronize()
        return buf.cpu().numpy().view(np.complex64).reshape(channels, samples)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import siggen
    uniq = min(channels, 4)
    base = [siggen.gen_shaped(samples, w["S"], w["M"], seed=seed + i, sigma=w["sigma"], freq=w["freq_max"] * 0.5) for i in range(uniq)]
    return np.stack([base[i % uniq] for i in range(channels)])


def run_reference(args, w):
    """--impl reference: the reference's CPU implementation of the path, all host threads, on a
    bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    try:
        import torch
        device_ok = torch.cuda.is_available()
    except Exception:
        device_ok = False
    ch = min(w["channels"], 512)
    n = min(w["samples"], 1_000_000) if w["channels"] > 1 else min(w["samples"], 8_000_000)
    iq = host_sample(w, ch, n, seed=1234, device_ok=device_ok)
    for _ in range(args.warmup):
        cpu_demod_rate(w, iq[: max(cores, 1)], cores)
    t_tot = 0.0
    kind = "port"
    for _ in range(args.steps):
        _, kind, dt = cpu_demod_rate(w, iq, cores)
        t_tot += dt
    ms = 1e3 * t_tot / args.steps
    val = ch * n / (ms * 1e-3) / 1e6
    sample = f"{ch} channels x {n} samples per step ({'oracle/_ref = unmodified psk_soft.cpp' if kind == 'reference' else 'oracle C port'}), one component per channel"
    line = {"impl": "reference", "metric": "Msamples/s demodulated", "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "packet_len": PACKET_LEN, "xdelta": XDELTA},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bank8psk", choices=sorted(WORKLOADS))
    ap.add_argument("--channels", type=int, default=0, help="override channels per GPU")
    ap.add_argument("--samples", type=int, default=0, help="override samples per channel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-channels", type=int, default=512)
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.channels:
        w["channels"] = args.channels
    if args.samples:
        w["samples"] = args.samples
    if args.impl == "reference":
        return run_reference(args, w)

    import numpy as np
    import torch
    import torch.distributed as dist
    import psk_soft_b200 as pk
    from psk_soft_b200 import binding as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the demod path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.cuda.current_device()
    full_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    nch, n = w["channels"], w["samples"]
    S, bpb = w["S"], bits_per_baud(w["M"])
    cap = n // S + 8
    iq = torch.empty((nch, n, 2), dtype=torch.float32, device="cuda")
    soft = torch.empty((nch, cap, 2), dtype=torch.float32, device="cuda")
    phase = torch.empty((nch, cap), dtype=torch.float32, device="cuda")
    sidx = torch.empty((nch, cap), dtype=torch.int16, device="cuda")
    bits = torch.empty((nch, cap * 3), dtype=torch.int16, device="cuda")
    pk.synth_fill(iq.data_ptr(), n, rank * nch, nch, n, seed=4, samplesPerBaud=S, constelationSize=w["M"],
                  sigma=w["sigma"], freq_max=w["freq_max"], pn_sigma=w["pn"], device=dev)
    torch.cuda.synchronize()
    bank = pk.Bank(nch, props_of(w), device=dev)
    stream = torch.cuda.ExternalStream(bank.stream, device=dev)

    def step():
        bank.process_raw(iq.data_ptr(), n, n, soft.data_ptr(), bits.data_ptr(), phase.data_ptr(), sidx.data_ptr(),
                         cap, cap * 3, xdelta=XDELTA, packet_len=PACKET_LEN, flags=B.FLAG_NO_SYNC, counts=False)

    def barrier():
        torch.cuda.synchronize()
        bank.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    bank.sync()
    bank.profile_read(reset=True)
    bank.profile_enable(True)
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    launches0 = bank.launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    launches = bank.launch_count - launches0
    time.sleep(0.15)
    sampler.stop()
    kern = bank.profile_read(reset=True)
    bank.profile_enable(False)
    stats = bank.stats()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    total_samples = world * nch * n
    value = total_samples / (ms_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (live CUDA-event time per launch) ----------------------
    peak, peak_src = measured_peak()
    roof = None
    if kern:
        K = n // S
        own = {"k_front": nch * (8 * n + 2 * K), "k_fused": algorithmic_bytes(w, nch, n),
               "k_chain_par": nch * K * 4, "k_chain_seq": nch * K * 4,
               "k_back_par": nch * K * (8 + 2 * bpb), "k_back": nch * K * (8 + 2 * bpb)}   # each kernel's own share of the algorithmic bytes
        dom = max(kern, key=lambda k: kern[k][0])
        ms_launch = kern[dom][0] / max(kern[dom][1], 1)
        launches_per_step = kern[dom][1] / args.steps
        abytes_step = algorithmic_bytes(w, nch, n)
        abytes_dom = own.get(dom, nch * K * 4)               # chain kernels: the 4-byte phase output
        achieved = abytes_dom / launches_per_step / (ms_launch * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes of the same kernel from the committed ncu --set full capture, scaled per launch
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if dom == "k_fused" and tr.get("workload") == args.workload:
                traffic = tr["dram_bytes_per_sample"] * nch * n / launches_per_step
        except (OSError, KeyError, ValueError):
            traffic = None
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes_dom / launches_per_step,
                "note": "achieved = this kernel's own algorithmic bytes (SURVEY 8d split per kernel) / its CUDA-event time per launch",
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kern.items()},
                "whole_path": {"algorithmic_bytes_per_step": abytes_step,
                               "achieved": abytes_step / (ms_step * 1e-3) / 1e9,
                               "frac": abytes_step / (ms_step * 1e-3) / 1e9 / peak}}

    # ---- e2e: same metric through the C ABI with HOST buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        per_ch = n * 8 + cap * (8 + 4 + 2 + 6)
        avail = psutil.virtual_memory().available
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        ech = max(1, min(nch, args.e2e_channels, int(0.25 * avail / max(local_world, 1) / per_ch)))
        h_iq = torch.empty((ech, n, 2), dtype=torch.float32, pin_memory=True)
        h_iq.copy_(iq[:ech])
        h_soft = torch.empty((ech, cap, 2), dtype=torch.float32, pin_memory=True)
        h_phase = torch.empty((ech, cap), dtype=torch.float32, pin_memory=True)
        h_sidx = torch.empty((ech, cap), dtype=torch.int16, pin_memory=True)
        h_bits = torch.empty((ech, cap * 3), dtype=torch.int16, pin_memory=True)
        ebank = pk.Bank(ech, props_of(w), device=dev)

        def estep():
            return ebank.process_raw(h_iq.data_ptr(), n, n, h_soft.data_ptr(), h_bits.data_ptr(), h_phase.data_ptr(),
                                     h_sidx.data_ptr(), cap, cap * 3, xdelta=XDELTA, packet_len=PACKET_LEN,
                                     flags=B.FLAG_HOST_BUFFERS)

        for _ in range(max(1, min(args.warmup, 2))):
            estep()
        barrier()
        esteps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(esteps):
            _, ns, nb = estep()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / esteps
        K = int(ns[0])
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * ech * n / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": ech * n * 8,
               "d2h_bytes_per_step": ech * K * (8 + 4 + 2 + 2 * bpb), "channels": ech, "samples": n,
               "ms_per_step": dt * 1e3, "note": "pskd_process with PSKD_FLAG_HOST_BUFFERS on pinned host memory; "
               f"{ech}-channel sub-bank of the workload (PCIe-bound, rate independent of bank size)"}
        del ebank

    # ---- cpu_baseline: the reference CPU demod on a bounded sample, rank 0, N=1 only --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, full_affinity)       # the reference gets every host core again
        cores = os.cpu_count() or 1
        cch = min(nch, 512)                      # ~20-30 core-seconds of reference CPU work
        cn = n if nch > 1 else min(n, 16_000_000)
        iq_h = iq[:cch, :cn].contiguous().cpu().numpy().view(np.complex64).reshape(cch, cn)
        rate, kind, secs = cpu_demod_rate(w, iq_h, cores)
        rate1, _, _ = cpu_demod_rate(w, iq_h[:1], 1)
        cpu = {"value": rate, "unit": "Msamples/s", "cores": cores, "kind": kind,
               "sample": f"first {cch} channels x {cn} samples of this workload, {secs:.2f} s wall on {cores} threads",
               "single_core_value": rate1}

    if rank == 0:
        clocks = sampler.summary(t_wall0, t_wall1)
        line = {"metric": "Msamples/s demodulated", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32/f64", "data": "synthetic",
                "config": {"workload": args.workload, "description": w["desc"], "channels_per_gpu": nch, "samples_per_channel": n,
                           "samplesPerBaud": S, "constelationSize": w["M"], "numAvg": w["A"], "phaseAvg": w["P"],
                           "differentialDecoding": w["D"], "packet_len": PACKET_LEN, "xdelta": XDELTA,
                           "l2": f"inputs ({nch * n * 8 / 1e9:.1f} GB per GPU) far larger than the 126 MB L2; no flush needed",
                           "parallelism": f"channels sharded over {world} GPU(s), no collective", "host_affinity": numa},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "chain": {k: stats[k] for k in ("spec_chunks", "spec_misses", "seq_channels", "wraps")}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
