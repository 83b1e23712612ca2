#!/usr/bin/env python
"""GPU tool: per-tile timeline (clock64 stamps of CTA 0) of the warp-specialised tensor-core BlazeBlock kernel.
Usage: tc_trace.py blk TR NSTG nsets nbuf esets [size] [batch]   (TR 0 = the library's default geometry; blk -1 = stem;
nbuf carries + 16 x work unit + 64 x issuers + 512 x placement as in hp_debug_set_tc)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpose_b200 import _lib  # noqa: E402
from hpose_b200.device import default_context  # noqa: E402
from hpose_b200.unified import pack_backbone, random_backbone  # noqa: E402

blk, TR, NSTG, nsets, nbuf, esets = [int(v) for v in sys.argv[1:7]]
size = int(sys.argv[7]) if len(sys.argv) > 7 else 96
B = int(sys.argv[8]) if len(sys.argv) > 8 else 4096
ctx = default_context()
lib = _lib.lib()
flat = pack_backbone(random_backbone(1234))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
H = size // 2 if blk < 2 else size // 4 if blk < 5 else size // 8 if blk < 11 else size // 16   # blk -1 = stem
if blk >= 0:
    _lib.check(lib.hp_debug_set_tc(ctx.handle, blk, TR, NSTG, 0, esets, nsets, nbuf))
else:
    _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, 0, nbuf, 0, nsets))
x = torch.rand((B, size, size, 3), device="cuda") * 2 - 1
NT = 40
trace = torch.zeros((NT, 12), dtype=torch.int64, device="cuda")
c = (24, 28, 32, 36, 42, 48, 56, 64, 72, 80, 88, 96, 96, 96, 96, 96)[blk] if blk >= 0 else 24
dst = torch.empty((B, H, H, c), device="cuda")
for rep in range(2):
    _lib.check(lib.hp_debug_tc_trace(ctx.handle, trace.data_ptr() if rep == 1 else None, NT))
    _lib.check(lib.hp_backbone_read_activation(ctx.handle, x.data_ptr(), B, size, size, blk, dst.data_ptr(), dst.numel(), None))
    torch.cuda.synchronize()
_lib.check(lib.hp_debug_tc_trace(ctx.handle, None, 0))
t = trace.cpu().numpy()
t0 = t[0, 0]
if os.environ.get("TC_RAW"):   # absolute stamps of the first tiles relative to slot 11 of tile 0 (kernel entry in the pixel-per-lane kernel)
    print("tile: slots 0..11 (clk after t[0][11])")
    for i in range(int(os.environ["TC_RAW"])):
        print(i, " ".join(f"{int(v - t[0, 11]):7d}" for v in t[i]))
print("per tile (clk): load = issue->full | dw0/dwL = full->set0/last set done | a_full = last set done->issuer sees last a_full | mma = ->d_full seen by epilogue")
print("tile   load   dw0   dwL a_full   mma   epi st_wait store | period  issuer: D-free wait (dempty->first a_full)")
for i in range(8, NT):
    per = t[i, 4] - t[i - 1, 4]
    print(f"{i:4d} {t[i,1]-t[i,0]:6d} {t[i,2]-t[i,1]:5d} {t[i,8]-t[i,1]:5d} {t[i,9]-t[i,8]:6d} {t[i,3]-t[i,9]:5d} {t[i,4]-t[i,3]:5d} {t[i,5]-t[i,4]:7d} {t[i,6]-t[i,5]:5d} | {per:6d}  {t[i,10]-t[i,11]:6d}")
