#!/usr/bin/env python
"""GPU tool: time the tensor-core BlazeBlock kernel (csrc/blocks_tc.cu) under different pipeline geometries
(hp_debug_set_tc: TR rows per lane, ring depth NSTG, band height BH, pipelines per CTA) with hp_backbone_profile.
Usage: tc_sweep.py [size] [batch];  writes gpurun_out/tc_sweep_<size>.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpose_b200 import _lib  # noqa: E402
from hpose_b200.device import default_context  # noqa: E402
from hpose_b200.unified import pack_backbone, random_backbone  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ctx = default_context()
lib = _lib.lib()
flat = pack_backbone(random_backbone(1234))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
x = torch.rand((B, size, size, 3), device="cuda") * 2 - 1
ms = np.zeros(18, dtype=np.float32)

# candidates (TR, NSTG, npipe, nsets); the band height follows from TR and the map size
# (TR, NSTG, npipe, nsets, nbuf): nbuf 0 = pipelined kernel; nbuf > 0 = warp-specialised kernel (npipe slot = epilogue warp sets,
# NSTG 0 = automatic, nbuf % 16 caps the ring of halo buffers, + 16 x work unit + 64 x issuer warps; band height / images per tile
# are chosen by the library)
CANDS = [(4, 0, 2, 2, 4 + 32 + 128), (4, 0, 1, 2, 4 + 32 + 128 + 512), (4, 0, 2, 2, 4 + 32 + 128 + 512), (4, 0, 1, 3, 4 + 32 + 128 + 512),
         (2, 0, 2, 3, 4 + 32 + 128), (2, 0, 1, 3, 4 + 32 + 128 + 512), (2, 0, 2, 2, 4 + 32 + 128 + 512), (2, 0, 2, 3, 4 + 32 + 128 + 512),
         (1, 0, 2, 3, 4), (1, 0, 1, 3, 4), (1, 0, 1, 2, 4), (1, 0, 2, 4, 4), (1, 0, 2, 3, 2), (1, 0, 2, 2, 4), (1, 0, 1, 4, 4),
         (1, 0, 2, 3, 4 + 16), (1, 0, 2, 2, 4 + 16)]   # TR 1: pixel-per-lane kernel
TC_BLOCKS = [int(b) for b in os.environ.get("TC_BLOCKS", "0,1,3,4,6,7,8,9,10,12").split(",")]
HS = {0: size // 2, 1: size // 2, 2: size // 4, 5: size // 8, 11: size // 16, 3: size // 4, 4: size // 4, 6: size // 8, 7: size // 8, 8: size // 8, 9: size // 8, 10: size // 8,
      12: size // 16, 13: size // 16, 14: size // 16, 15: size // 16}


def band_height(H, W, TR):
    strips = -(-H // TR)
    mx = max(1, 128 // W)
    bands = -(-strips // mx)
    return -(-strips // bands) * TR


def run():
    _lib.check(lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, size, size, 5, ms.ctypes.data))
    return ms.copy()


run()  # warm-up (module load, attribute setup)
base = run()
print("default:", " ".join(f"b{b}={base[1 + b]:.4f}" for b in TC_BLOCKS), flush=True)
res = {"default": {f"block{b}": float(base[1 + b]) for b in TC_BLOCKS}}
for b in TC_BLOCKS:
    _lib.check(lib.hp_debug_set_tc(ctx.handle, b, -1, 0, 0, 0, 0, 0))
t = run()
res["cuda_core"] = {f"block{b}": float(t[1 + b]) for b in TC_BLOCKS}
print("cuda-core kernel:", " ".join(f"b{b}={t[1 + b]:.4f}" for b in TC_BLOCKS), flush=True)
for b in TC_BLOCKS:
    for cand in CANDS:
        TR, NSTG, npipe, nsets, nbuf = cand
        BH = band_height(HS[b], HS[b], TR)
        _lib.check(lib.hp_debug_set_tc(ctx.handle, b, TR, NSTG, BH, npipe, nsets, nbuf))
        try:
            t = run()
            res.setdefault(f"block{b}", {})[str(cand)] = float(t[1 + b])
            print(f"block{b} TR={TR} NSTG={NSTG} BH={BH} npipe={npipe} nsets={nsets} nbuf={nbuf % 16} unit={max(1, (nbuf // 16) % 4)} issuers={max(1, (nbuf // 64) % 8)} placed={nbuf // 512}: {t[1 + b]:.4f} ms", flush=True)
        except Exception as e:  # configuration not instantiated / does not fit
            print(f"block{b} {cand} failed: {str(e)[:120]}", flush=True)
    _lib.check(lib.hp_debug_set_tc(ctx.handle, b, -1, 0, 0, 0, 0, 0))
for b in TC_BLOCKS:
    _lib.check(lib.hp_debug_set_tc(ctx.handle, b, 0, 0, 0, 0, 0, 0))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"tc_sweep_{size}.json"), "w"), indent=1)
