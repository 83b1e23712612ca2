import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
sys.path.insert(0, ".")
from hpose_b200 import _lib
from hpose_b200.device import default_context
from hpose_b200.unified import pack_backbone, random_backbone
size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = 4096
ctx = default_context(); lib = _lib.lib()
flat = pack_backbone(random_backbone(1234))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
x = torch.rand((B, size, size, 3), device="cuda") * 2 - 1
ms = np.zeros(18, dtype=np.float32)
def run():
    for _ in range(2):
        _lib.check(lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, size, size, 5, ms.ctypes.data))
    return ms.copy()
base = run()
print("default:", " ".join(f"b{b}={base[1+b]:.4f}" for b in (0,1,3,4)), flush=True)
HS = {0: size // 2, 1: size // 2, 3: size // 4, 4: size // 4}
def band_height(H, W, TR):
    strips = -(-H // TR); mx = max(1, 128 // W); bands = -(-strips // mx)
    return -(-strips // bands) * TR
# (TR, nsets, esets, unit, issuers, placed)
INST = [(4,2,2,1,1,0),(4,3,2,1,1,0),(2,2,2,1,1,0),(2,3,2,1,1,0),(4,2,1,1,1,0),(2,3,1,1,1,0),(2,2,2,2,1,0),(2,3,2,2,1,0),(2,4,1,2,1,0),(4,2,2,2,1,0),
        (4,3,2,2,1,0),(4,2,2,2,2,0),(4,3,2,2,2,0),(2,3,2,2,2,0),(2,2,2,2,2,0),(2,3,2,1,2,0),(4,2,1,2,2,1),(4,2,2,2,2,1),(4,3,1,2,2,1),(2,3,1,2,2,1),(2,2,2,2,2,1),(2,3,2,2,2,1)]
for b in (3, 4, 0, 1):
    res = []
    for (TR, nsets, esets, unit, niss, place) in INST:
        for ring in (4, 3):
            nbuf = ring + 16 * unit + 64 * niss + 512 * place
            BH = band_height(HS[b], HS[b], TR)
            try:
                _lib.check(lib.hp_debug_set_tc(ctx.handle, b, TR, 0, BH, esets, nsets, nbuf))
                t = run()
                res.append((float(t[1 + b]), (TR, nsets, esets, unit, niss, place, ring)))
            except Exception as e:
                pass
    _lib.check(lib.hp_debug_set_tc(ctx.handle, b, 0, 0, 0, 0, 0, 0))
    res.sort()
    for t, c in res[:8]:
        print(f"block{b} TR,nsets,esets,unit,niss,place,ring={c}: {t:.4f} ms", flush=True)
