"""Chain kernel (blocks_chain.cu) check on a GPU: per-block agreement with the per-block kernels and the naive kernels,
then per-layer timings with the chain on / off.  Usage: python tools/chain_check.py [size] [batch] [stage...]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpose_b200 import _lib  # noqa: E402
from hpose_b200.device import default_context  # noqa: E402
from hpose_b200.unified import pack_backbone, random_backbone  # noqa: E402

BLOCKS = [(24, 24, 1), (24, 28, 1), (28, 32, 2), (32, 36, 1), (36, 42, 1), (42, 48, 2), (48, 56, 1), (56, 64, 1),
          (64, 72, 1), (72, 80, 1), (80, 88, 1), (88, 96, 2), (96, 96, 1), (96, 96, 1), (96, 96, 1), (96, 96, 1)]


def sizes(S):
    h = -(-S // 2)
    out = []
    for cin, cout, s in BLOCKS:
        h = -(-h // s)
        out.append((h, cout))
    return out


def read_act(ctx, x, blk, shape):
    dst = torch.empty(shape, dtype=torch.float32, device=x.device)
    B, H, W, _ = x.shape
    _lib.check(_lib.lib().hp_backbone_read_activation(ctx.handle, x.data_ptr(), B, H, W, blk, dst.data_ptr(), dst.numel(), ctx.stream_ptr()))
    torch.cuda.synchronize()
    return dst


def chain_status(ctx, where):
    st = np.zeros(8, np.uint32)
    _lib.check(_lib.lib().hp_debug_chain_status(ctx.handle, st.ctypes.data))
    if st[0]:
        names = {1: "tile_full", 2: "epi", 3: "a_empty", 4: "d_full", 5: "w_full", 6: "a_full", 7: "w_empty", 8: "tile_done"}
        print(f"CHAIN WATCHDOG at {where}: barrier {names.get(int(st[1]), st[1])} parity {st[2]} thread {st[3]} (warp {st[3] // 32}) CTA {st[4]} step {st[5]}", flush=True)
    return bool(st[0])


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def parity(ctx, S, B):
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(S)
    x = torch.rand((B, S, S, 3), generator=g, device="cuda") * 2 - 1
    sz = sizes(S)
    worst = 0.0
    for blk in range(5, 16):
        h, c = sz[blk]
        _lib.check(L.hp_debug_set_chain(ctx.handle, 0, 0, 0))
        ref = read_act(ctx, x, blk, (B, h, h, c))
        ctx.set_impl(_lib.HP_IMPL_NAIVE)
        nv = read_act(ctx, x, blk, (B, h, h, c))
        ctx.set_impl(_lib.HP_IMPL_FAST)
        _lib.check(L.hp_debug_set_chain(ctx.handle, 2, 0, 0))
        got = read_act(ctx, x, blk, (B, h, h, c))
        if chain_status(ctx, f"block {blk}"):
            return 1.0
        e1, e2 = rel(got, ref), rel(got, nv)
        worst = max(worst, e2)
        print(f"size {S} B {B} block {blk:2d} ({h}x{h}x{c}): chain vs per-block {e1:.2e}  vs naive {e2:.2e}  (per-block vs naive {rel(ref, nv):.2e})", flush=True)
    rep = np.zeros(128, np.int32)
    _lib.check(L.hp_debug_tile_report(ctx.handle, rep.ctypes.data))
    read_act(ctx, x, 15, (B, sz[15][0], sz[15][0], 96))
    print("tile report blocks 6, 12:", rep.reshape(16, 8)[6].tolist(), rep.reshape(16, 8)[12].tolist())
    _lib.check(L.hp_debug_tile_report(ctx.handle, None))
    return worst


def timing(ctx, S, B, iters=5):
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((B, S, S, 3), generator=g, device="cuda") * 2 - 1
    for mode, nsets, niss in ((0, 0, 0), (1, 0, 0), (2, 0, 0), (2, 4, 2), (6, 0, 0), (6, 4, 2)):   # + 4: 3xTF32 instead of split fp16
        _lib.check(L.hp_debug_set_chain(ctx.handle, mode, nsets, niss))
        per = np.zeros(18, np.float32)
        for _ in range(2):
            _lib.check(L.hp_backbone_profile(ctx.handle, x.data_ptr(), B, S, S, iters, per.ctypes.data))
        print(f"size {S} B {B} chain mode {mode} nsets {nsets} niss {niss}: backbone {per[:17].sum():.3f} ms | blocks 6-10 {per[7:12].sum():.3f} | "
              f"block 11 {per[12]:.3f} | blocks 12-15 {per[13:17].sum():.3f} | stem {per[0]:.3f} | 0-5 {per[1:7].sum():.3f}", flush=True)
    _lib.check(L.hp_debug_set_chain(ctx.handle, 2, 0, 0))


def trace(ctx, S, B, steps=40):
    """Clock stamps of CTA 0 (8 per step).  The run up to block 10 contains only chain 6-10; the run up to block 15 contains both
    chains, chain 12-15 last (it overwrites the first steps of the buffer: only those are printed)."""
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((B, S, S, 3), generator=g, device="cuda") * 2 - 1
    sz = sizes(S)
    for blk_last, name, nblk in ((10, "chain 6-10", 5), (11, "chain 6-10 + tail block 11", 6), (15, "chain 12-15", 4)):
        buf = torch.zeros(steps * 8, dtype=torch.int64, device="cuda")
        h, c = sz[blk_last]
        read_act(ctx, x, blk_last, (B, h, h, c))
        _lib.check(L.hp_debug_tc_trace(ctx.handle, C.c_void_p(buf.data_ptr()), steps))
        read_act(ctx, x, blk_last, (B, h, h, c))
        _lib.check(L.hp_debug_tc_trace(ctx.handle, None, 0))
        t = buf.cpu().numpy().reshape(steps, 8)
        t0 = t[0, 0]
        print(f"--- {name}: per step (clk since first step ready): ready | dw done | d_full seen | epi done | first a_full | mma issued | store issued | store read")
        for i in range(steps):
            if t[i, 0] == 0 or (i > 0 and t[i, 0] < t[i - 1, 0]):
                break
            row = [int(v - t0) if v else -1 for v in t[i]]
            print(f"{i:3d} blk+{i % nblk} " + " ".join(f"{v:8d}" for v in row) + f"   | dw {row[1]-row[0]:6d} mma-tail {row[2]-row[1]:6d} epi {row[3]-row[2]:6d}")


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    stages = sys.argv[3:] or ["parity"]
    ctx = default_context()
    w = random_backbone(seed=1234, bias_scale=0.05)
    flat = pack_backbone(w)
    _lib.check(_lib.lib().hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
    if "parity" in stages:
        worst = parity(ctx, S, B)
        print("WORST", worst, "OK" if worst < 2e-5 else "FAIL")
    if "timing" in stages:
        timing(ctx, S, B)
    if "trace" in stages:
        trace(ctx, S, B)
