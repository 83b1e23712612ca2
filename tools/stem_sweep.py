"""GPU tool: tensor-core stem, agreement with the naive kernel and timing of the pipeline geometries.
Usage: python tools/stem_sweep.py [size] [batch]
cfg = (BH, nbuf, nout, sets): sets = gather sets + 16 x issuers (+ 128: issuers on SM sub-partition 3, first generation only;
+ 256: first-generation kernel, lane <-> column of a strip of rows)"""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from hpose_b200 import _lib  # noqa: E402
from hpose_b200.device import default_context  # noqa: E402
from hpose_b200.unified import pack_backbone, random_backbone  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ctx = default_context()
lib = _lib.lib()
flat = pack_backbone(random_backbone(1234))
_lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))


def read_stem(x):
    b, h, w, _ = x.shape
    dst = torch.empty((b, h // 2, w // 2, 24), device="cuda")
    _lib.check(lib.hp_backbone_read_activation(ctx.handle, x.data_ptr(), b, h, w, -1, dst.data_ptr(), dst.numel(), None))
    torch.cuda.synchronize()
    return dst


# agreement with the one-thread-per-output kernel on small batches of every supported shape class
for s, b in ((96, 5), (88, 3), (128, 3), (64, 2), (120, 2), (100, 2), (104, 2), (160, 1)):
    xs = torch.rand((b, s, s, 3), device="cuda") * 2 - 1
    ctx.set_impl(_lib.HP_IMPL_NAIVE)
    want = read_stem(xs)
    ctx.set_impl(_lib.HP_IMPL_FAST)
    for cfg in ((0, 0, 0, 0), (0, 0, 0, 256), (0, 2, 2, 3 + 32), (4, 3, 2, 2 + 32), (0, 0, 0, 2 + 64), (0, 0, 0, 3 + 48 + 512)):
        try:
            _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, *cfg))
            got = read_stem(xs)
            err = float((got - want).abs().max() / want.abs().max())
            print(f"size {s} batch {b} cfg {cfg}: rel err {err:.2e}", "OK" if err < 2e-5 else "MISMATCH", flush=True)
        except Exception as e:
            print(f"size {s} cfg {cfg} failed: {str(e)[:120]}", flush=True)
_lib.check(lib.hp_debug_set_stem_tc(ctx.handle, 0, 0, 0, 0))

x = torch.rand((B, size, size, 3), device="cuda") * 2 - 1
ms = np.zeros(18, dtype=np.float32)
cfgs = [(0, 0, 0, 0), (0, 0, 0, 256), (0, 0, 0, 2 + 32), (0, 0, 0, 4 + 32), (0, 0, 0, 2 + 16), (0, 0, 0, 2 + 64)]
for ne in (128, 512):
    for ns, ni in ((2, 2), (3, 2), (4, 2), (3, 3), (4, 3), (2, 4), (3, 4), (4, 4)):
        cfgs.append((0, 0, 0, ns + 16 * ni + ne))
cfgs += [(0, 3, 2, 0), (0, 2, 2, 0), (0, 4, 3, 0), (4, 0, 0, 0), (-1, 0, 0, 0)]
for cfg in cfgs:
    try:
        _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, *cfg))
        for _ in range(2):
            _lib.check(lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, size, size, 5, ms.ctypes.data))
        print(cfg, "sets", cfg[3] % 16, "issuers", (cfg[3] // 16) % 8, "epilogue sets", 3 if cfg[3] & 512 else 2 if cfg[3] & 128 else 1,
              "gen1" if cfg[3] & 256 else "flat", f"stem {ms[0]:.4f} ms", flush=True)
    except Exception as e:
        print(cfg, "failed", str(e)[:100], flush=True)
_lib.check(lib.hp_debug_set_stem_tc(ctx.handle, 0, 0, 0, 0))
