"""Chain kernel timing under the HP_CHAIN_EXP switches (1 = no depthwise loads / math, 2 = no epilogue shared-memory traffic,
4 = no MMAs; the outputs are garbage, only the time matters).  Usage: HP_CHAIN_EXP=n python tools/chain_exp.py [size] [batch]"""
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
for _ in range(3):
    _lib.check(L.hp_backbone_profile(ctx.handle, x.data_ptr(), B, S, S, 5, per.ctypes.data))
print(f"HP_CHAIN_EXP={os.environ.get('HP_CHAIN_EXP', '0')} {os.environ.get('HP_CHAIN_VAR', '')} size {S} B {B}: backbone {per[:17].sum():.3f} ms | "
      f"chain 6-11 {per[7:13].sum():.4f} | chain 12-15 {per[13:17].sum():.4f}", flush=True)
