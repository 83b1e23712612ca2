"""Per-layer backbone times (hp_backbone_profile, CUDA events around 5 launches of each kernel) and the fraction of the HBM
roofline.  Usage: python tools/layer_times.py [size] [batch] [label]"""
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

S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
label = sys.argv[3] if len(sys.argv) > 3 else os.environ.get("HPOSE_LIB_SUFFIX", "")
BLOCKS = [(24, 24, 1), (24, 28, 1), (28, 32, 2), (32, 36, 1), (36, 42, 1), (42, 48, 2), (48, 56, 1), (56, 64, 1),
          (64, 72, 1), (72, 80, 1), (80, 88, 1), (88, 96, 2), (96, 96, 1), (96, 96, 1), (96, 96, 1), (96, 96, 1)]
peak = 6534.5e9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass
ctx = default_context()
L = _lib.lib()
flat = pack_backbone(random_backbone(seed=1234, bias_scale=0.05))
_lib.check(L.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
x = torch.rand((B, S, S, 3), device="cuda") * 2 - 1
per = np.zeros(18, np.float32)
for _ in range(3):
    _lib.check(L.hp_backbone_profile(ctx.handle, x.data_ptr(), B, S, S, 5, per.ctypes.data))
h = -(-S // 2)
byts = [(S * S * 3 + h * h * 24) * 4]
for cin, cout, s in BLOCKS:
    ho = -(-h // s)
    byts.append((h * h * cin + ho * ho * cout) * 4)
    h = ho
names = ["stem"] + [f"b{i}" for i in range(16)]
out, i = [], 0
while i < 17:
    j = i + 1
    while j < 17 and per[j] == 0.0 and per[i] > 0.0 and i >= 7:
        j += 1                                   # a chain kernel books its time on its first block
    by = sum(byts[i:j]) * B
    nm = names[i] if j == i + 1 else f"{names[i]}-{names[j - 1][1:]}"
    out.append(f"{nm} {per[i]:.3f} ({by / (per[i] * 1e-3) / peak:.2f})")
    i = j
tot = float(per[:17].sum())
print(f"[{label}] {S}x{S} B {B}: backbone {tot:.3f} ms = {sum(byts) * B / (tot * 1e-3) / peak:.3f} of the HBM roofline | " + " ".join(out), flush=True)
