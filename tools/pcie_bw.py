#!/usr/bin/env python
"""GPU tool: pinned host <-> device copy bandwidth for the e2e batch sizes (is the serving loop PCIe / host-memory bound?).

    python tools/pcie_bw.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_bw.py

Under torchrun every rank drives its own GPU; the copies are timed (a) with ALL ranks copying at the same time and (b) rank by
rank with the others idle, so the loss of per-GPU bandwidth under concurrency -- the shared host root complex / memory --
is measured, not asserted.  Rank 0 prints one JSON line.  Sizes: the headline step's H2D (4096 x 96 x 96 x 3 uint8 = 113 MB)
and padded D2H (59 MB)."""
import json
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n_in, n_out = 4096 * 96 * 96 * 3, 58998784
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


def measure():
    t_in = timed(lambda: d_in.copy_(h_in, non_blocking=True))
    t_out = timed(lambda: h_out.copy_(d_out, non_blocking=True))
    t_both = timed(both)
    return [n_in / t_in / 1e6, n_out / t_out / 1e6, (n_in + n_out) / t_both / 1e6, t_both]


res = {"world": world, "h2d_MB": n_in / 1e6, "d2h_MB": n_out / 1e6}
conc = torch.tensor(measure(), dtype=torch.float64, device="cuda")
if world > 1:
    allc = [torch.empty_like(conc) for _ in range(world)]
    dist.all_gather(allc, conc)
    # alone: one rank at a time
    alone = torch.zeros(4, dtype=torch.float64, device="cuda")
    for r in range(world):
        dist.barrier()
        if r == rank:
            e = []
            for fn in (lambda: d_in.copy_(h_in, non_blocking=True), lambda: h_out.copy_(d_out, non_blocking=True), both):
                fn(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record(); torch.cuda.synchronize()
                e.append(e0.elapsed_time(e1) / 10)
            alone = torch.tensor([n_in / e[0] / 1e6, n_out / e[1] / 1e6, (n_in + n_out) / e[2] / 1e6, e[2]], dtype=torch.float64, device="cuda")
        dist.barrier()
    alla = [torch.empty_like(alone) for _ in range(world)]
    dist.all_gather(alla, alone)
    if rank == 0:
        res["concurrent_per_rank_GBps"] = [{"h2d": float(c[0]), "d2h": float(c[1]), "h2d+d2h": float(c[2]), "pair_ms": float(c[3])} for c in allc]
        res["alone_per_rank_GBps"] = [{"h2d": float(c[0]), "d2h": float(c[1]), "h2d+d2h": float(c[2]), "pair_ms": float(c[3])} for c in alla]
        res["aggregate_concurrent_GBps"] = {"h2d": float(sum(c[0] for c in allc)), "d2h": float(sum(c[1] for c in allc)),
                                            "h2d+d2h": float(sum(c[2] for c in allc))}
        res["per_gpu_bandwidth_kept_under_concurrency"] = float(sum(c[2] for c in allc) / sum(c[2] for c in alla))
        res["step_copy_ms_concurrent_max"] = float(max(c[3] for c in allc))
else:
    res["alone_GBps"] = {"h2d": float(conc[0]), "d2h": float(conc[1]), "h2d+d2h": float(conc[2]), "pair_ms": float(conc[3])}
try:
    res["cpu_count"] = os.cpu_count()
    res["affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]
except Exception:
    pass
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
