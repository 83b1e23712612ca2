#!/usr/bin/env python
"""GPU tool: sweep BlazeBlock tile shapes (hp_debug_set_tile) and time every kernel (hp_backbone_profile).
Writes gpurun_out/tile_sweep_<size>.json; used to tune choose_tile() in csrc/backbone.cu."""
import ctypes as C
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

BLOCKS = [(24, 24, 1), (24, 28, 1), (28, 32, 2), (32, 36, 1), (36, 42, 1), (42, 48, 2), (48, 56, 1), (56, 64, 1),
          (64, 72, 1), (72, 80, 1), (80, 88, 1), (88, 96, 2), (96, 96, 1), (96, 96, 1), (96, 96, 1), (96, 96, 1)]


def cd(a, b):
    return -(-a // b)


def au(v, a=32):
    return cd(v, a) * a


def valid(Hout, Wout, S, cin, cout, TH, TW, IMGS, nbuf, MT):
    CINP, COUTP = (cin + 3) & ~3, (cout + 3) & ~3
    NG, C4 = COUTP // 4, CINP // 4
    TP = TH * TW * IMGS
    if TH < 1 or TW < 1 or TH > Hout or TW > Wout or TP < 32 or TP > 512:
        return False
    PG = cd(TP, MT)
    thr = cd(PG * NG, 32) * 32
    if thr > 512 or thr < C4:
        return False
    IH, IW = (TH - 1) * S + 3, (TW - 1) * S + 3
    DWS = CINP if C4 & 1 else CINP + 4
    off = 32
    off = au(off + CINP * COUTP + COUTP + 10 * CINP)
    off = au(off + 2 * IMGS * TH * cd(TW, 4) + PG * MT)
    off = au(off + max(PG * MT * DWS, TP * COUTP))
    off += nbuf * au(IMGS * IH * IW * CINP) + au((4 * S + 3) * CINP)
    return off * 4 <= 227 * 1024


def candidates(Hout, Wout, S, cin, cout):
    out = set()
    ths = {Hout, cd(Hout, 2), cd(Hout, 3), cd(Hout, 4), cd(Hout, 6), 2, 3, 4, 6, 8, 12, 16}
    tws = {Wout, cd(Wout, 2), cd(Wout, 3), cd(Wout, 4), 4, 6, 8, 12, 16, 24, 32}
    for TH in ths:
        for TW in tws:
            for IMGS in (1, 2, 3, 4, 6, 8):
                if IMGS > 1 and (TH != Hout or TW != Wout):
                    continue
                for nbuf in (1, 2):
                    for MT in (4, 8):
                        if valid(Hout, Wout, S, cin, cout, TH, TW, IMGS, nbuf, MT):
                            out.add((TH, TW, IMGS, nbuf, MT))
    return sorted(out)


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    ctx = default_context()
    lib = _lib.lib()
    flat = pack_backbone(random_backbone(1234))
    _lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
    x = torch.rand((B, size, size, 3), device="cuda") * 2 - 1
    report = np.zeros(128, np.int32)
    lib.hp_debug_tile_report(ctx.handle, report.ctypes.data)
    per = np.zeros(18, np.float32)
    # heuristic baseline
    _lib.check(lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, size, size, 3, per.ctypes.data))
    _lib.check(lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, size, size, 5, per.ctypes.data))
    base = {"per_layer_ms": per.tolist(), "tiles": report.reshape(16, 8).tolist()}
    print("heuristic:", [round(float(v), 4) for v in per], "total", float(per[:17].sum()))
    h = cd(size, 2)
    cands = []
    for (cin, cout, s) in BLOCKS:
        ho = cd(h, s)
        cands.append(candidates(ho, ho, s, cin, cout))
        h = ho
    results = [dict() for _ in range(16)]
    rounds = max(len(c) for c in cands)
    for r in range(rounds):
        for i in range(16):
            c = cands[i][r % len(cands[i])]
            lib.hp_debug_set_tile(ctx.handle, i, *c)
        rc = lib.hp_backbone_profile(ctx.handle, x.data_ptr(), B, size, size, 3, per.ctypes.data)
        if rc != 0:
            print("round", r, "failed:", lib.hp_last_error().decode())
            torch.cuda.synchronize()
            continue
        for i in range(16):
            c = cands[i][r % len(cands[i])]
            results[i][",".join(map(str, c))] = float(per[1 + i])
    for i in range(16):
        lib.hp_debug_set_tile(ctx.handle, i, 0, 0, 0, 0, 0)
    best_total = float(per[0])
    summary = []
    for i in range(16):
        ranked = sorted(results[i].items(), key=lambda kv: kv[1])
        summary.append({"block": i, "heuristic_ms": base["per_layer_ms"][1 + i], "heuristic_tile": base["tiles"][i],
                        "best": ranked[:6], "worst": ranked[-2:]})
        best_total += ranked[0][1] if ranked else base["per_layer_ms"][1 + i]
        print(i, BLOCKS[i], "heur", round(base["per_layer_ms"][1 + i], 4), base["tiles"][i], "best", ranked[:4])
    print("stem", base["per_layer_ms"][0], "sum-of-best", best_total, "heuristic total", sum(base["per_layer_ms"][:17]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"tile_sweep_{size}.json"), "w") as f:
        json.dump({"size": size, "batch": B, "baseline": base, "summary": summary, "all": results}, f)


if __name__ == "__main__":
    main()
