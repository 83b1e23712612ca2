import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from hpose_b200 import _lib
from hpose_b200.device import default_context
from hpose_b200.unified import pack_backbone, random_backbone
ctx = default_context(); lib = _lib.lib()
flat = pack_backbone(random_backbone(7, bias_scale=0.1))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
B, S = 5, 96
x = torch.rand((B, S, S, 3), device="cuda") * 2 - 1
A = lib.hp_num_anchors(S, S)
cls = torch.empty((B, A), device="cuda"); loc = torch.empty((B, A, 16), device="cuda")
_lib.check(lib.hp_backbone_forward(ctx.handle, x.data_ptr(), B, S, S, None, None, cls.data_ptr(), loc.data_ptr(), None))
torch.cuda.synchronize()
print("ok", float(cls.abs().mean()))
