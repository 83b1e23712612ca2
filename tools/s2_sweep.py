"""Stride-2 block kernel (blocks 2, 5): band height R x input buffers sweep.  Usage: python tools/s2_sweep.py [size] [batch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpose_b200 import _lib  # noqa: E402
from hpose_b200.device import default_context  # noqa: E402
from hpose_b200.unified import pack_backbone, random_backbone  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ctx = default_context()
L = _lib.lib()
flat = pack_backbone(random_backbone(seed=1234, bias_scale=0.05))
_lib.check(L.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
x = torch.rand((B, S, S, 3), device="cuda") * 2 - 1
per = np.zeros(18, np.float32)


def run():
    for _ in range(2):
        _lib.check(L.hp_backbone_profile(ctx.handle, x.data_ptr(), B, S, S, 5, per.ctypes.data))
    return per.copy()


base = run()
print(f"size {S} B {B} default: b2 {base[3]:.4f} b5 {base[6]:.4f}", flush=True)
for blk, Wo in ((2, -(-S // 4)), (5, -(-S // 8))):
    for R in range(1, min(Wo, 128 // Wo) + 1):
        for nbuf in (2, 3, 4):
            try:
                _lib.check(L.hp_debug_set_tc(ctx.handle, blk, 1, 0, R, 2, 2, nbuf))
                t = run()
                print(f"block{blk} R={R} nbuf<={nbuf}: {t[1 + blk]:.4f} ms", flush=True)
            except Exception as e:
                print(f"block{blk} R={R} nbuf={nbuf} failed: {str(e)[:90]}", flush=True)
    _lib.check(L.hp_debug_set_tc(ctx.handle, blk, 0, 0, 0, 0, 0, 0))
