"""GPU tool: timing experiments of the flat tensor-core stem (env HP_STEM_EXP, see StemFlatParams::exp_).  Usage: HP_STEM_EXP=n python tools/stem_exp.py [size] [batch] [sets]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from hpose_b200 import _lib
from hpose_b200.device import default_context
from hpose_b200.unified import pack_backbone, random_backbone
size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sets = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = default_context(); lib = _lib.lib()
flat = pack_backbone(random_backbone(1234))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
x = torch.rand((B, size, size, 3), device="cuda") * 2 - 1
ms = np.zeros(18, dtype=np.float32)
_lib.check(lib.hp_debug_set_stem_tc(ctx.handle, 0, 0, 0, sets))
for _ in range(2):
    _lib.check(lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, size, size, 5, ms.ctypes.data))
print("HP_STEM_EXP", os.environ.get("HP_STEM_EXP", "0"), "sets", sets, f"stem {ms[0]:.4f} ms", flush=True)
