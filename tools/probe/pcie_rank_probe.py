"""What caps the end-to-end (host-buffer) rate when several GPUs of one box are fed at once?
Run under torchrun with N ranks (one per GPU):  H2D alone, D2H alone, and both at once (D2H moving a quarter of the
bytes, like the demod outputs), per rank and summed over the ranks, with pinned buffers first-touched by the rank that
uses them.  Prints ONE JSON line on rank 0:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe/pcie_rank_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True); h_out.fill_(1)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())                      # the slowest rank ends the step
    return (n / dt / 1e9 if h2d else 0.0), (n / 4 / dt / 1e9 if d2h else 0.0)


run(True, True, 2)
a = run(True, False)[0]
b = run(False, True)[1]
c, d = run(True, True)
if rank == 0:
    aff = sorted(os.sched_getaffinity(0))
    print(json.dumps({"n_gpus": world, "per_rank_GBps": {"h2d_alone": a, "d2h_alone": b, "duplex_h2d": c, "duplex_d2h": d},
                      "box_GBps": {"h2d_alone": a * world, "d2h_alone": b * world, "duplex_h2d": c * world, "duplex_d2h": d * world},
                      "host_cpus": len(aff), "note": "1 GiB pinned H2D (+ 256 MiB D2H) per rank per step, slowest rank"}), flush=True)
if world > 1:
    dist.destroy_process_group()
