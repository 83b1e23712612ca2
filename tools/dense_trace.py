#!/usr/bin/env python
"""GPU tool: time one Dense / 1x1-conv layer (hp_debug_dense) and print the per-tile clock stamps of CTA 0 of the
tensor-core kernel.  Usage: dense_trace.py M K N [n2] [act]   (n2 > 0: fused narrow second layer)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpose_b200 import _lib  # noqa: E402
from hpose_b200.device import default_context  # noqa: E402

M, K, N = [int(v) for v in sys.argv[1:4]]
n2 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ctx = default_context()
lib = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((M, K), generator=g, device="cuda")
W = torch.randn((K, N), generator=g, device="cuda") / K ** 0.5
b = torch.randn((N,), generator=g, device="cuda")
W2 = torch.randn((N, max(n2, 1)), generator=g, device="cuda") / N ** 0.5
b2 = torch.randn((max(n2, 1),), generator=g, device="cuda")
y = torch.empty((M, n2 if n2 else N), device="cuda")
NT = 24
trace = torch.zeros((NT, 12), dtype=torch.int64, device="cuda")


def run():
    _lib.check(lib.hp_debug_dense(ctx.handle, x.data_ptr(), M, K, W.data_ptr(), b.data_ptr(), N, act, y.data_ptr(),
                                  W2.data_ptr() if n2 else None, b2.data_ptr() if n2 else None, n2, 0, None))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = (M * K + M * (n2 if n2 else N)) * 4 / 1e9
print(f"M {M} K {K} N {N} n2 {n2}: {ms * 1e3:.1f} us, {gb / ms * 1e3:.0f} GB/s algorithmic, {-(-M // 128) / 148:.1f} tiles per SM")
_lib.check(lib.hp_debug_tc_trace(ctx.handle, trace.data_ptr(), NT))
run()
torch.cuda.synchronize()
_lib.check(lib.hp_debug_tc_trace(ctx.handle, None, 0))
t = trace.cpu().numpy()
print("tile: slots 0..11 (clk after kernel entry): 0 load issued, 1 full seen, 2 set0 done, 3 d_full seen, 4 D released, 5 written, 7 MMAs issued, 8 last set done, 9/10 last/first a_full, 11 D free")
for i in range(NT):
    if t[i, 0] == 0:
        break
    print(i, " ".join(f"{int(v - t[0, 11]):7d}" for v in t[i]))
