#!/usr/bin/env python
"""Small, fixed workload for ncu: two backbone forwards (stem + 16 BlazeBlocks + detector heads) at
B crops of size x size with random-init weights.  Usage: profile_target.py [size] [batch] [unified]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpose_b200 import _lib  # noqa: E402
from hpose_b200.device import default_context  # noqa: E402
from hpose_b200.unified import pack_backbone, random_backbone  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ctx = default_context()
lib = _lib.lib()
flat = pack_backbone(random_backbone(1234))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
# optional geometry override for the tensor-core block kernel: TC_GEOM="blk:TR,NSTG,BH,npipe,nsets,nbuf;blk:..."
for item in filter(None, os.environ.get("TC_GEOM", "").split(";")):
    blk, vals = item.split(":")
    _lib.check(lib.hp_debug_set_tc(ctx.handle, int(blk), *[int(v) for v in vals.split(",")]))
if os.environ.get("STEM_CFG"):   # tensor-core stem geometry: "BH,nbuf,nout,sets" as in hp_debug_set_stem_tc
    _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, *[int(v) for v in os.environ["STEM_CFG"].split(",")]))
x = torch.rand((B, size, size, 3), device="cuda") * 2 - 1
A = lib.hp_num_anchors(size, size)
cls = torch.empty((B, A), device="cuda")
loc = torch.empty((B, A, 16), device="cuda")
for _ in range(2):
    _lib.check(lib.hp_backbone_forward(ctx.handle, x.data_ptr(), B, size, size, None, None, cls.data_ptr(), loc.data_ptr(), None))
torch.cuda.synchronize()
print("ok", float(cls.abs().mean()))
