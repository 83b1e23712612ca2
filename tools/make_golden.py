#!/usr/bin/env python
"""Regenerates tests/golden/* from the read-only reference checkout (/root/reference).

Run in the build container only (the GPU box has no /root/reference).  Nothing here is
product code; it reads the reference's *artefacts* (trained .h5 checkpoints and .npz feature
datasets) with the repo's own mini-HDF5 reader and stores small fixtures:

  unified_graph.json        model_config of BlazePoser/UnifiedModels/reg1-stoqa9pt-reg2-hrchr82r-selected.h5
                            with the inlined Constant initializers removed (graph only)
  unified_weights.npz       its weights; the detector part is float16-exact and is
                            stored as float16, nested head weights as float32
  heads/<id>.h5             three small trained regressor checkpoints, byte-identical artefacts
  kat_heads.npz / .json     dataset slices + float64 oracle predictions + whole-dataset MAE known answers
  unified_kat.npz           float64 oracle outputs of the unified graph on a seeded U(-1,1) image pair
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "head-pose-estimation-model_b200"))
import h5lite  # noqa: E402
from oracle.keras_graph import KerasGraph, normalise_weight_names, to_torch  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
UNIFIED = f"{REF}/BlazePoser/UnifiedModels/reg1-stoqa9pt-reg2-hrchr82r-selected.h5"
HEADS = {
    "stoqa9pt": f"{REF}/Model-88/Trained-Models-88/stoqa9pt.h5",
    "12uei1sn": f"{REF}/Model-88/Trained-Models-88/12uei1sn.h5",
    "hrchr82r": f"{REF}/Model-96/Trained-Models-96-ReshapedInput-NoFlatten/hrchr82r.h5",
}
DATASETS = {
    "stoqa9pt": ["BIWI_Test_Enlarged_features_88_0.7_1.npz", "AFLW2000_Enlarged_features_88_0.7_1.npz",
                 "BIWI_Train_Enlarged_features_88_0.7_1.npz"],
    "12uei1sn": ["BIWI_Test_Enlarged_features_88_0.7_1.npz", "AFLW2000_Enlarged_features_88_0.7_1.npz"],
    "hrchr82r": ["AFLW2000_features_96_0.7_1.npz"],
}


def strip_constants(obj):
    if isinstance(obj, dict):
        if obj.get("class_name") == "Constant" and "config" in obj:
            return {"class_name": "Constant", "config": {"value": 0.0}}
        return {k: strip_constants(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [strip_constants(v) for v in obj]
    return obj


def main():
    os.makedirs(os.path.join(OUT, "heads"), exist_ok=True)
    f = h5lite.H5File(UNIFIED)
    cfg = f.model_config()
    graph = strip_constants(cfg)
    with open(os.path.join(OUT, "unified_graph.json"), "w") as fh:
        json.dump(graph, fh, separators=(",", ":"))
    w = normalise_weight_names(f.weights())
    packed = {}
    for k, v in w.items():
        v16 = v.astype(np.float16)
        if np.array_equal(v16.astype(np.float32), v):
            packed[k] = v16
        else:
            packed[k] = v
    np.savez_compressed(os.path.join(OUT, "unified_weights.npz"), **packed)
    n16 = sum(1 for v in packed.values() if v.dtype == np.float16)
    print(f"unified: {len(packed)} tensors ({n16} float16-exact), "
          f"{sum(v.size for v in packed.values())} params")

    # unified KAT: two seeded images through the float64 oracle
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, size=(2, 128, 128, 3)).astype(np.float32)
    g = KerasGraph(graph, to_torch(w, torch.float64))
    taps = {}
    with torch.no_grad():
        outs = g(torch.tensor(x, dtype=torch.float64), taps=taps)
    np.savez_compressed(os.path.join(OUT, "unified_kat.npz"), x=x,
                        cls16=outs[0].numpy(), cls8=outs[1].numpy(), loc16=outs[2].numpy(),
                        loc8=outs[3].numpy(), pose16=outs[4].numpy(), pose8=outs[5].numpy(),
                        tap16=taps["re_lu_10"].numpy(), tap8=taps["re_lu_15"].numpy(),
                        stem=taps["conv2d"].numpy()[:, :8, :8], blk2=taps["re_lu_2"].numpy()[:, :4, :4])

    kat = {}
    slices = {}
    for hid, path in HEADS.items():
        shutil.copyfile(path, os.path.join(OUT, "heads", f"{hid}.h5"))
        os.chmod(os.path.join(OUT, "heads", f"{hid}.h5"), 0o644)
        hf = h5lite.H5File(path)
        hw = normalise_weight_names(hf.weights())
        hg = KerasGraph(hf.model_config(), to_torch(hw, torch.float64))
        first_kernel = sorted(k for k in hw if k.endswith("kernel"))[0]
        kat[hid] = {"params": int(sum(v.size for v in hw.values())),
                    "first_kernel": first_kernel,
                    "first_kernel_sha16": hashlib.sha256(hw[first_kernel].astype("<f4").tobytes()).hexdigest()[:16],
                    "datasets": {}}
        for ds in DATASETS[hid]:
            d = np.load(f"{REF}/FeatureMaps-Datasets/{ds}")
            feats, poses = d["features"], d["poses"]
            c = feats.shape[1]
            with torch.no_grad():
                pred = hg(torch.tensor(feats.reshape(-1, 1, 1, c), dtype=torch.float64)).numpy().reshape(-1, 3)
            mae = np.abs(pred - poses).mean(axis=0)
            mse = ((pred - poses) ** 2).mean(axis=0)
            kat[hid]["datasets"][ds] = {"n": int(feats.shape[0]), "mae_ypr": mae.tolist(), "mae_avg": float(mae.mean()),
                                       "mse_avg": float(mse.mean()), "row0": pred[0].tolist()}
            key = f"{hid}|{ds}"
            slices[f"{key}|features"] = feats[:256]
            slices[f"{key}|poses"] = poses[:256]
            slices[f"{key}|pred64"] = pred[:256]
            print(hid, ds, "MAE", mae.mean(), "row0", pred[0])
    with open(os.path.join(OUT, "kat_heads.json"), "w") as fh:
        json.dump(kat, fh, indent=1)
    np.savez_compressed(os.path.join(OUT, "kat_heads.npz"), **slices)


if __name__ == "__main__":
    main()
