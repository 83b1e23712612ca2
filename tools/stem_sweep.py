import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from hpose_b200 import _lib
from hpose_b200.device import default_context
from hpose_b200.unified import pack_backbone, random_backbone
ctx = default_context(); lib = _lib.lib()
flat = pack_backbone(random_backbone(1234))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
B = 4096
x = torch.rand((B, 96, 96, 3), device="cuda") * 2 - 1
ms = np.zeros(18, dtype=np.float32)
# cfg[3] = gather sets + 16 x issuers + 128 x (issuers placed on SM sub-partition 3)
for cfg in [(0,0,0,0), (0,4,2,2+32), (0,4,2,2+32+128), (0,4,2,3+32+128), (0,4,2,2+16+128), (0,4,2,3+32), (-1,0,0,0)]:
    try:
        _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, *cfg))
        for _ in range(2):
            _lib.check(lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, 96, 96, 5, ms.ctypes.data))
        print(cfg, "nsets", cfg[3] % 16, "issuers", (cfg[3] // 16) % 8, "placed", cfg[3] // 128, f"stem {ms[0]:.4f} ms", flush=True)
    except Exception as e:
        print(cfg, "failed", str(e)[:100])
