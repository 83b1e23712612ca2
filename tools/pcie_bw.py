#!/usr/bin/env python
"""GPU tool: pinned host <-> device copy bandwidth for the e2e batch sizes (is the serving loop PCIe-bound?)."""
import torch

n_in, n_out = 4096 * 96 * 96 * 3, 58998784
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t = timed(lambda: d_in.copy_(h_in, non_blocking=True))
print(f"H2D {n_in / 1e6:.0f} MB: {t:.3f} ms, {n_in / t / 1e6:.1f} GB/s")
t = timed(lambda: h_out.copy_(d_out, non_blocking=True))
print(f"D2H {n_out / 1e6:.0f} MB: {t:.3f} ms, {n_out / t / 1e6:.1f} GB/s")


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


t = timed(both)
print(f"H2D + D2H concurrently: {t:.3f} ms per pair")
