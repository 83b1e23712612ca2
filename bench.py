#!/usr/bin/env python
"""Headline benchmark: pose inferences/sec on synthetic 96x96 crops (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement (oracle)

One step = one pass of the unified hot path over one batch of 4096 crops per GPU:
  BlazeFace backbone -> taps 12x12x88 / 6x6x96 -> detector heads + pose heads
  (Model-88 shipped shape 88->64(softsign)->3 on the 88-channel tap, se_transformer_regr_head(96) on
  the 96-channel tap) -> threshold + anchor decode + NMS + pose lookup.
`value`  : crops/s with the float32 inputs already resident in HBM (device-timed, max over ranks).
`e2e`    : the same metric through the Python facade with HOST buffers: pinned uint8 BGR crops are
           copied host->device every step, pre-processed on the GPU, and counts/boxes/keypoints/scores/poses
           are read back to pinned host memory every step.
`roofline`: per-kernel CUDA-event timings measured live (hp_backbone_profile) against the measured HBM
           copy bandwidth in MEASURED_PEAKS.json; algorithmic bytes per SURVEY 8(d) / DESIGN.md; `traffic` is the
           dram__bytes_read + write of the same kernel from the committed ncu --set full capture (profiles/).
`dtype`  : f32 -- all tensors and accumulators are fp32; the pointwise / stem GEMMs run on the tensor cores as
           split-precision products (3xTF32, stem: split fp16) with fp32-level error.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCKS = [(24, 24, 1), (24, 28, 1), (28, 32, 2), (32, 36, 1), (36, 42, 1), (42, 48, 2), (48, 56, 1), (56, 64, 1),
          (64, 72, 1), (72, 80, 1), (80, 88, 1), (88, 96, 2), (96, 96, 1), (96, 96, 1), (96, 96, 1), (96, 96, 1)]
METRIC = "pose inferences/sec at 96x96 crops"
UNIT = "crops/s"


def layer_table(size):
    """[(name, algorithmic bytes per crop, flops per crop)] -- stem + one read of the input and one write of the
    output per BlazeBlock, fp32, weights excluded."""
    rows = []
    h = -(-size // 2)
    rows.append(("stem", (size * size * 3 + h * h * 24) * 4, 2 * h * h * 24 * 75))
    for i, (cin, cout, s) in enumerate(BLOCKS):
        ho = -(-h // s)
        rows.append((f"block{i}", (h * h * cin + ho * ho * cout) * 4, 2 * ho * ho * (9 * cin + cin * cout)))
        h = ho
    return rows


def ncu_traffic(kernel, batch, size):
    """DRAM bytes per launch of one backbone kernel from the committed ncu --set full capture (profiles/): None when
    the capture was taken at another batch / size."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for root, _, files in os.walk(pdir):
        for f in sorted(files):
            if f.startswith("ncu_traffic") and f.endswith(".json"):
                try:
                    with open(os.path.join(root, f)) as fh:
                        d = json.load(fh)
                    e = d.get(kernel)
                    if e and int(e.get("batch", 0)) == batch and int(e.get("size", 0)) == size:
                        best = float(e["dram_read_bytes"]) + float(e["dram_write_bytes"])
                except Exception:
                    pass
    return best


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_model(size):
    from hpose_b200 import keras_spec as K, train_88
    from hpose_b200.attention_model import se_transformer_regr_head
    from hpose_b200.unified import UnifiedModel, random_backbone
    K.reset_names(); K.set_seed(1234)
    head16 = train_88.create_model()
    K.reset_names()
    head8 = se_transformer_regr_head(input_channels=96)
    return UnifiedModel(random_backbone(seed=1234), head16, head8)


def cpu_reference_run(size, sample, reps, threads):
    """Times the CPU restatement (oracle) of the same path on `sample` crops; returns crops/s (best of reps)."""
    import torch
    from oracle import postproc as opp
    from oracle.keras_graph import KerasGraph, to_torch
    torch.set_num_threads(threads)
    model = build_model(size)
    w = {}
    from hpose_b200.unified import unpack_backbone
    w.update(unpack_backbone(model.backbone_flat))
    for name, head in ((model.head16_name, model.head16), (model.head8_name, model.head8)):
        for k, v in head.program.unpack(head._flat).items():
            w[f"{name}/{k}"] = v
    g = KerasGraph(model.config(), to_torch(w, torch.float32))
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(sample, size, size, 3), dtype=np.uint8)
    anchors = opp.blazeface_anchors(size)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        x = ((img[..., ::-1].astype(np.float64) / 255.0).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
        with torch.no_grad():
            o = [t.numpy() for t in g(torch.from_numpy(x))]
        for i in range(sample):
            cls = np.concatenate([o[0][i, :, 0], o[1][i, :, 0]])
            loc = np.concatenate([o[2][i], o[3][i]])
            opp.detect_postprocess(cls, loc, o[4][i], o[5][i], anchors, 0.4, 0.3, input_size=size)
        best = min(best, time.perf_counter() - t0)
    return sample / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_sample
    times = []
    for _ in range(args.warmup):
        cpu_reference_run(args.size, sample, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cps, dt = cpu_reference_run(args.size, sample, 1, threads)
        times.append(dt)
    total = time.perf_counter() - t0
    value = sample * args.steps / sum(times)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"unified head-pose path, {args.size}x{args.size} uint8 crops, batch {args.batch}/GPU "
                                   f"(reference arm times a bounded sample of {sample} crops per step)",
                       "batch_per_gpu": args.batch, "input": f"{args.size}x{args.size}x3"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} crops/step x {args.steps} steps; restated CPU path (TensorFlow/Keras "
                                       "unavailable): torch-CPU fp32 graph + numpy decode/NMS"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": total}
    print(json.dumps(line))



def merged_layers(size, per, B, peak):
    """Per-kernel rows: a fused chain (blocks 6-10 / 12-15 as one persistent kernel) books its time on its first block and the
    others report 0 ms -- they are merged into one row whose algorithmic bytes are the SUM of the blocks' SURVEY 8(d) figures
    (the per-unit accounting the roofline fraction is defined on); `fused_min_bytes` is what the fused kernel has to move
    (input of the first block + output of the last)."""
    rows = layer_table(size)
    out = []
    for (name, byts, flops), t_ms in zip(rows, per[:17]):
        t_ms = float(t_ms)
        if t_ms <= 0.0 and out:
            o = out[-1]
            o["blocks"].append(name)
            o["bytes"] += byts
            o["flops"] += flops
            continue
        out.append({"blocks": [name], "bytes": byts, "flops": flops, "ms": t_ms})
    h = -(-size // 2)
    dims = {}
    for i, (cin, cout, st) in enumerate(BLOCKS):
        ho = -(-h // st)
        dims[f"block{i}"] = (h * h * cin * 4, ho * ho * cout * 4)
        h = ho
    layers = []
    for o in out:
        nm = o["blocks"][0] if len(o["blocks"]) == 1 else f"blocks{o['blocks'][0][5:]}-{o['blocks'][-1][5:]}"
        gbs = o["bytes"] * B / (o["ms"] * 1e-3) / 1e9 if o["ms"] > 0 else 0.0
        row = {"kernel": nm, "ms": o["ms"], "algorithmic_GBps": gbs, "frac": gbs / peak,
               "tflops": o["flops"] * B / (o["ms"] * 1e-3) / 1e12 if o["ms"] > 0 else 0.0, "algorithmic_bytes_per_crop": o["bytes"]}
        if len(o["blocks"]) > 1:
            row["fused_min_bytes_per_crop"] = dims[o["blocks"][0]][0] + dims[o["blocks"][-1]][1]
        layers.append(row)
    return layers


def backbone_profile(ctx, x, B, S, peak, iters=5):
    from hpose_b200 import _lib
    per = np.zeros(18, np.float32)
    _lib.check(_lib.lib().hp_backbone_profile(ctx.handle, x.data_ptr(), B, S, S, iters, per.ctypes.data))
    layers = merged_layers(S, per, B, peak)
    rows = layer_table(S)
    bb_ms = float(per[:17].sum())
    bb_bytes = sum(r[1] for r in rows)
    bb_flops = sum(r[2] for r in rows)
    gbs = bb_bytes * B / (bb_ms * 1e-3) / 1e9
    return {"ms": bb_ms, "achieved": gbs, "frac": gbs / peak, "bytes_per_crop": bb_bytes, "flop_per_crop": bb_flops,
            "fp32_tflops": bb_flops * B / (bb_ms * 1e-3) / 1e12, "det_heads_ms": float(per[17])}, layers


def time_steps(fn, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extra_configs(args, ctx, peak):
    """BASELINE.json configs 2, 3 and 5 (single GPU), each with its own backbone roofline fraction, the NMS stress input of
    SURVEY 8(d) config 3, and a seconds-long sustained run of the headline config with the SM clock seen under load."""
    import torch
    from hpose_b200 import _lib
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    dev = ctx.torch_device
    out = {}

    def unified(S, B, steps):
        det = blazeFaceDetector(model=build_model(S), inputSize=S)
        g = torch.Generator(device=dev).manual_seed(S * 100003 + B)
        x = torch.rand((B, S, S, 3), generator=g, device=dev) * 2 - 1
        res = {}
        ms = time_steps(lambda: res.update(o=det.detect_device(x, args.max_faces, float_input=True, out=res.get("o"))), steps)
        bb, _ = backbone_profile(ctx, x, B, S, peak, iters=3)
        del x, res
        torch.cuda.empty_cache()
        return {"input": f"{S}x{S}x3", "batch": B, "ms_per_step": ms, "crops_per_s": B / (ms * 1e-3), "backbone_ms": bb["ms"],
                "backbone_frac_of_hbm_roofline": bb["frac"], "backbone_GBps": bb["achieved"]}

    # config 2: Model-88 variant, 88x88 crops, batch 1024 (11x11x88 tap -> 88->64->3 head, 6x6x96 tap -> attention head)
    out["config2_model88_88x88_b1024"] = unified(88, 1024, 20)
    # config 3: BlazeFace 128x128 front detector full path, batch 4096
    c3 = unified(128, 4096, 8)
    rng = np.random.default_rng(7)                      # NMS stress input (bypasses the backbone): SURVEY 8(d) config 3
    Bn, A = 4096, 896
    cls = torch.from_numpy(rng.normal(0, 2, (Bn, A)).astype(np.float32)).to(dev)
    loc = np.zeros((Bn, A, 16), np.float32)
    loc[..., :2] = rng.uniform(-8, 8, (Bn, A, 2))
    loc[..., 2:4] = rng.uniform(16, 64, (Bn, A, 2))
    loc[..., 4:] = rng.uniform(-20, 20, (Bn, A, 12))
    loc = torch.from_numpy(loc).to(dev)
    p16 = torch.zeros((Bn, 16, 16, 3), device=dev)
    p8 = torch.zeros((Bn, 8, 8, 3), device=dev)
    cnt = torch.empty((Bn,), dtype=torch.int32, device=dev)
    anc = torch.empty((Bn, 100), dtype=torch.int32, device=dev)
    boxes = torch.empty((Bn, 100, 4), dtype=torch.float64, device=dev)
    kps = torch.empty((Bn, 100, 12), dtype=torch.float64, device=dev)
    sc = torch.empty((Bn, 100), dtype=torch.float32, device=dev)
    po = torch.empty((Bn, 100, 3), dtype=torch.float32, device=dev)
    thr = float(np.float32(np.log(0.4 / 0.6)))

    def nms():
        _lib.check(_lib.lib().hp_decode_nms(ctx.handle, cls.data_ptr(), loc.data_ptr(), p16.data_ptr(), p8.data_ptr(), Bn, 128, 128, thr,
                                            float(np.float32(0.3)), 100, cnt.data_ptr(), anc.data_ptr(), boxes.data_ptr(), kps.data_ptr(),
                                            sc.data_ptr(), po.data_ptr(), ctx.stream_ptr()))
    ms = time_steps(nms, 10)
    alg = Bn * (3584 + 57344 + 3840 + 8000)             # read cls + loc + poses, write <= 100 faces
    c3["nms_stress_seed7"] = {"ms": ms, "images_per_s": Bn / (ms * 1e-3), "mean_faces_kept": float(cnt.float().mean().item()),
                              "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_roofline": alg / (ms * 1e-3) / 1e9 / peak}
    out["config3_detector_128x128_b4096"] = c3
    del cls, loc, p16, p8, cnt, anc, boxes, kps, sc, po
    # config 5: unified detector + pose end to end, batch sweep at the reference's real input size
    out["config5_unified_128x128_batch_sweep"] = [unified(128, b, max(3, min(20, 32768 // b))) for b in (256, 512, 1024, 2048, 4096, 8192, 16384)]
    return out


def sustained_run(args, det, x, seconds=3.0):
    """The headline step in a loop of >= `seconds` s: throughput and SM clock under sustained load (the 20-step headline run
    lasts 80 ms and sits at the maximum boost clock)."""
    import torch
    res = {}
    step = lambda: res.update(o=det.detect_device(x, args.max_faces, float_input=True, out=res.get("o")))
    ms1 = time_steps(step, 5)
    n = int(max(20, seconds * 1e3 / ms1))
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    ms = time_steps(step, n, warmup=0)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    return {"steps": n, "seconds": n * ms * 1e-3, "ms_per_step": ms, "crops_per_s": x.shape[0] / (ms * 1e-3), "clocks": sampler.summary()}


def latency_run(args, ctx):
    """The reference's real usage (one webcam frame per call, blazeFaceDetectorH5.py:392-444): detectFaces on 640x480 uint8
    frames through the single-call path (one H2D, one CUDA-graph launch of resize + unified graph + decode + NMS, one D2H);
    small batches through the same entry.  Wall-clock per call, host side included."""
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    det = blazeFaceDetector(model=build_model(128), inputSize=128)
    rng = np.random.default_rng(5)
    out = {}
    for B in (1, 8, 64):
        frames = rng.integers(0, 256, size=(B, 480, 640, 3), dtype=np.uint8)
        for _ in range(5):
            det._detect_packed(frames, args.max_faces, graph=True)
        ts = []
        for _ in range(200 if B == 1 else 60):
            t0 = time.perf_counter()
            det._detect_packed(frames, args.max_faces, graph=True)
            ts.append(time.perf_counter() - t0)
        ts = np.sort(np.array(ts)) * 1e3
        nog = []
        for _ in range(30):
            t0 = time.perf_counter()
            det._detect_packed(frames, args.max_faces, graph=False)
            nog.append(time.perf_counter() - t0)
        out[f"batch{B}"] = {"frame": "640x480x3 uint8 -> 128x128", "p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)),
                            "frames_per_s_p50": B / (float(np.percentile(ts, 50)) * 1e-3),
                            "p50_ms_without_graph": float(np.percentile(np.array(nog) * 1e3, 50))}
    return out


def trained_e2e(args, ctx, steps, world=1):
    """End to end with the TRAINED detector + shipped pose heads (tests/golden, the reference's own weights) on synthetic
    frames: few anchors fire, so the packed result is a few bytes per crop instead of the 14.4 KB of the padded form."""
    import torch
    gold = os.path.join(ROOT, "tests", "golden")
    if not os.path.exists(os.path.join(gold, "unified_weights.npz")):
        return None
    from hpose_b200 import keras_spec as K
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel
    with np.load(os.path.join(gold, "unified_weights.npz")) as z:
        w = {k: z[k].astype(np.float32) for k in z.files}
    u = UnifiedModel(w, K.load_model(os.path.join(gold, "heads", "stoqa9pt.h5")), K.load_model(os.path.join(gold, "heads", "hrchr82r.h5")))
    det = blazeFaceDetector(model=u, inputSize=args.size)
    B, S = args.batch, args.size
    rng = np.random.default_rng(11)
    host = torch.from_numpy(rng.integers(0, 256, size=(B, S, S, 3), dtype=np.uint8)).pin_memory()

    def batches(nb):
        for _ in range(nb):
            yield host
    total = 0
    for r in det.detect_stream(batches(3), args.max_faces, packed=True):
        total = r["total"]
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in det.detect_stream(batches(steps), args.max_faces, packed=True):
        total = r["total"]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=ctx.torch_device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    hdr = (4 + B) * 4
    return {"value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "h2d_bytes_per_step": int(host.numel()),
            "d2h_bytes_per_step": int(hdr + total * 152), "faces_per_step": int(total),
            "weights": "trained BlazeFace-front detector + shipped stoqa9pt / hrchr82r heads (tests/golden)", "api": "detect_stream(packed=True)"}


def train_leg(args, ctx, rank, world):
    """BASELINE config 4: train_96 head (96 -> 64 -> ... -> 3, Adam lr 2.8e-4, l2 1e-5) on synthetic BIWI-shaped feature maps,
    data-parallel over the ranks with the gradient all-reduce inside the captured step (hp_head_train_run).  Weak scaling:
    128 rows per GPU (global 128 N); strong: global batch 128 (the reference's batch) split over the ranks."""
    import torch
    import torch.distributed as dist
    from hpose_b200 import keras_spec as K, train_96
    from hpose_b200.parallel import DataParallel
    dev = ctx.torch_device
    N = 131072
    rng = np.random.default_rng(4)
    x = (np.maximum(rng.normal(0, 0.55, (N, 96)), 0) * (rng.random((N, 96)) < 0.31)).astype(np.float32).reshape(N, 1, 1, 96)
    y = rng.normal([15, -6, -1], [27.5, 27, 12.5], (N, 3)).astype(np.float32).reshape(N, 1, 1, 3)
    xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    dp = None
    if world > 1:
        dp = DataParallel()
        dp.init_gradient_comm()
        out_p2p = bool(dp.p2p)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    out = {"model": "train_96.create_model(num_filters=64, dropout 0, l2 1e-5), Adam lr 2.8e-4", "items": N, "world": world}
    if world > 1:
        out["gradient_exchange"] = ("fused all-reduce + optimizer kernel over NVLink peer memory (CUDA IPC inboxes, rank-ordered sums)"
                                    if out_p2p else "ncclAllReduce + optimizer kernel (peer memory not available)")

    def make():
        train_96.config.update(num_filters=64, dropout_rate=0.0, regularizer_rate=1e-5, optimizer="adam")
        K.reset_names(); K.set_seed(3)
        m = train_96.create_model()
        m.optimizer.learning_rate = 2.8e-4
        return m

    def flat(m):
        return np.concatenate([v.reshape(-1) for v in m.get_weights_dict().values()])

    for label, gb in (("per_gpu_batch_128", 128 * world), ("global_batch_128", 128)):
        m = make()
        steps = min(2000, N // gb)
        with torch.cuda.stream(side):
            m.train_run_device(xt, yt, None, 0, gb, 20, rank, world, seed=1, graph=True)          # warm-up + capture
            side.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            sums = m.train_run_device(xt, yt, None, 0, gb, steps, rank, world, seed=1, graph=True)
            e1.record(side)
            side.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        r = {"global_batch": gb, "rows_per_gpu": gb // world, "steps": steps, "us_per_step": 1e3 * ms / steps, "steps_per_s": steps / (ms * 1e-3),
             "samples_per_s": steps * gb / (ms * 1e-3), "mean_loss": sums[0] / (steps * gb)}
        w = torch.from_numpy(flat(m)).to(dev)
        if world > 1:
            gathered = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(gathered, w)
            r["ranks_identical"] = bool(all(torch.equal(gathered[0], g) for g in gathered[1:]))
            # the same global batches on one GPU (this rank alone, no communicator involved)
            single = make()
            with torch.cuda.stream(side):
                single.train_run_device(xt, yt, None, 0, gb, 20, 0, 1, seed=1, graph=True)
                single.train_run_device(xt, yt, None, 0, gb, steps, 0, 1, seed=1, graph=True)
                side.synchronize()
            ref = flat(single)
            r["max_rel_dev_vs_single_gpu"] = float(np.abs(flat(m) - ref).max() / max(1.0, np.abs(ref).max()))
            # cost of the all-reduce inside the step: the same steps without it (each rank on its own shard, weights diverge)
            solo = make()
            with torch.cuda.stream(side):
                solo.train_run_device(xt, yt, None, 0, gb // world, 20, 0, 1, seed=1, graph=True)
                side.synchronize()
                dist.barrier()
                e0.record(side)
                solo.train_run_device(xt, yt, None, 0, gb // world, steps, 0, 1, seed=1, graph=True)
                e1.record(side)
                side.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r["us_per_step_without_allreduce"] = 1e3 * float(t.item()) / steps
            r["allreduce_us"] = r["us_per_step"] - r["us_per_step_without_allreduce"]
            if out_p2p:
                # the same steps with ncclAllReduce + a separate optimizer kernel (the library baseline of the fused exchange)
                from hpose_b200 import _lib
                _lib.check(_lib.lib().hp_debug_set_p2p(ctx.handle, 0))
                mn = make()
                with torch.cuda.stream(side):
                    mn.train_run_device(xt, yt, None, 0, gb, 20, rank, world, seed=1, graph=True)
                    side.synchronize()
                    dist.barrier()
                    e0.record(side)
                    mn.train_run_device(xt, yt, None, 0, gb, steps, rank, world, seed=1, graph=True)
                    e1.record(side)
                    side.synchronize()
                _lib.check(_lib.lib().hp_debug_set_p2p(ctx.handle, 1))
                t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                r["us_per_step_nccl"] = 1e3 * float(t.item()) / steps
                r["allreduce_us_nccl"] = r["us_per_step_nccl"] - r["us_per_step_without_allreduce"]
                r["max_rel_dev_p2p_vs_nccl"] = float(np.abs(flat(m) - flat(mn)).max() / max(1.0, np.abs(flat(mn)).max()))
        else:
            # host-driven baseline of round 1: one hp_head_train_step call per step, loss read back every step
            m2 = make()
            idx = torch.arange(gb, device=dev)
            xb, yb = xt[:gb].contiguous(), yt[:gb].contiguous()
            for _ in range(20):
                m2.train_on_device(xb, yb, seed=1)
            t0 = time.perf_counter()
            for _ in range(300):
                m2.train_on_device(xb, yb, seed=1)
            torch.cuda.synchronize()
            r["us_per_step_host_driven_with_loss_readback"] = 1e6 * (time.perf_counter() - t0) / 300
        out[label] = r
        if world == 1:
            break
    torch.cuda.current_stream(dev).wait_stream(side)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="crops per GPU per step")
    ap.add_argument("--size", type=int, default=96)
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--max-faces", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (other BASELINE configs, sustained run, latency, training)")
    ap.add_argument("--e2e-chunks", type=int, default=1,
                    help="slices per batch in the end-to-end serving loop (detect_stream chunks); measured at batch 4096: 1 -> 799k, 2 -> 762k, "
                         "4 -> 659k crops/s (smaller launches cost more than the shorter pipeline fill saves)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from hpose_b200 import _lib
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.device import default_context
    B, S = args.batch, args.size
    model = build_model(S)
    det = blazeFaceDetector(model=model, inputSize=S)
    ctx = default_context()
    dev = ctx.torch_device

    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    u8 = torch.randint(0, 256, (B, S, S, 3), generator=gen, device=dev, dtype=torch.uint8)
    x = det._preprocess_device(u8)        # float32 inputs resident in HBM for the device-timed region
    host_u8 = torch.empty((B, S, S, 3), dtype=torch.uint8).pin_memory()
    host_u8.copy_(u8.cpu())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return det.detect_device(x, args.max_faces, float_input=True)

    host_out = {}

    def step_e2e():
        d_u8 = host_u8.to(dev, non_blocking=True)
        out = det.detect_device(d_u8, args.max_faces)
        for k in ("count", "boxes", "keypoints", "scores", "poses"):
            if k not in host_out:
                host_out[k] = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
            host_out[k].copy_(out[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    for _ in range(args.warmup):
        out = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end-to-end through the facade with host buffers: blazeFaceDetector.detect_stream copies every step's pinned
    # uint8 batch host->device and every step's results device->host, overlapped with the kernels of the neighbouring steps
    for _ in range(2):
        step_e2e()
    barrier()
    e_steps = max(3, min(args.steps, 50))   # the same K steps as the device-resident measurement (pipeline fill / drain included)

    def batches(nb):
        for _ in range(nb):
            yield host_u8

    for res in det.detect_stream(batches(3), args.max_faces, chunks=args.e2e_chunks):
        host_out = res
    barrier()
    e0.record()
    for res in det.detect_stream(batches(e_steps), args.max_faces, chunks=args.e2e_chunks):
        host_out = res
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)
    t = torch.tensor([ems], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ems = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=3)
    e2e_value = world * B * e_steps / (ems * 1e-3)
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())

    if rank == 0:
        # ---- per-kernel roofline (CUDA events on the launch stream, each kernel repeated back to back)
        peak, peak_src = measured_peaks()
        bb, layers = backbone_profile(ctx, x, B, S, peak)
        bb_ms = bb["ms"]
        dom = max(layers, key=lambda r: r["ms"])
        try:
            fma_scalar = ctx.fma_peak_tflops(False)
            fma_packed = ctx.fma_peak_tflops(True)
        except Exception:
            fma_scalar = fma_packed = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"unified head-pose path (backbone + detector heads + Model-88 head 88->64->3 + "
                                       f"se_transformer_regr_head(96) + decode/NMS/pose lookup), {S}x{S} crops, "
                                       f"batch {B} per GPU, random-init weights seed 1234",
                           "batch_per_gpu": B, "global_batch": B * world, "input": f"{S}x{S}x3", "max_faces": args.max_faces,
                           "parallelism": f"dp{world} (batch sharded, no collective)",
                           "l2_policy": f"inputs larger than L2 ({B * S * S * 3 * 4 / 1e6:.0f} MB fp32 per step vs 126 MB)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_u8.numel()), "d2h_bytes_per_step": int(d2h),
                        "steps": e_steps, "ms_per_step": ems / e_steps,
                        "api": "blazeFaceDetector.detect_stream(pinned uint8 BGR host batches) -> pinned host count/boxes/keypoints/scores/"
                               "poses; H2D and D2H of neighbouring steps overlap the kernels (3 rotating buffer sets, 2 batches in flight)", "chunks": args.e2e_chunks},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["algorithmic_GBps"], "peak": peak,
                             "unit": "GB/s", "frac": dom["frac"], "traffic": ncu_traffic(dom["kernel"], B, S),
                             "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_crop"] * B, "peak_source": peak_src,
                             "ms_per_launch": dom["ms"], "share_of_backbone": dom["ms"] / bb_ms,
                             "fused_min_bytes_per_launch": dom.get("fused_min_bytes_per_crop", dom["algorithmic_bytes_per_crop"]) * B},
                "roofline_backbone": {"bound": "hbm", "achieved": bb["achieved"], "peak": peak, "unit": "GB/s", "frac": bb["frac"],
                                      "bytes_per_crop": bb["bytes_per_crop"], "flop_per_crop": bb["flop_per_crop"], "ms": bb_ms,
                                      "fp32_tflops": bb["fp32_tflops"],
                                      "fma_peak_tflops_measured": {"scalar_ffma": fma_scalar, "packed_f32x2": fma_packed},
                                      "det_heads_ms": bb["det_heads_ms"]},
                "layers": layers}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = max(args.cpu_sample, 1024)        # ~10 s of CPU work: a quarter of the GPU batch, twice
            cps, dt = cpu_reference_run(S, n_cpu, 2, threads)
            line["cpu_baseline"] = {"value": cps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_cpu} crops, best of 2; restated CPU path (TensorFlow/Keras "
                                              "unavailable): torch-CPU fp32 graph + numpy decode/NMS"}
        if world == 1 and not args.no_extras:
            line["sustained"] = sustained_run(args, det, x)
            line["latency"] = latency_run(args, ctx)
            line["configs"] = extra_configs(args, ctx, peak)
    train = e2e_tr = None
    if not args.no_extras:
        del x, u8, out
        torch.cuda.empty_cache()
        e2e_tr = trained_e2e(args, ctx, e_steps, world)    # every rank: host <-> device copies of all GPUs at the same time
        train = train_leg(args, ctx, rank, world)          # every rank takes part (gradient exchange)
    if rank == 0:
        line["e2e_trained_weights"] = e2e_tr
        line["train"] = train
        print(json.dumps(line, default=float))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
