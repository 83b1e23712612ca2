import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
from hpose_b200.device import default_context
from hpose_b200 import keras_spec as K
from hpose_b200.unified import UnifiedModel
ctx = default_context()
B, S = 4096, 96
def trained():
    gold = '/root/repo/tests/golden'
    with np.load(os.path.join(gold, "unified_weights.npz")) as z:
        w = {k: z[k].astype(np.float32) for k in z.files}
    return UnifiedModel(w, K.load_model(os.path.join(gold, "heads", "stoqa9pt.h5")), K.load_model(os.path.join(gold, "heads", "hrchr82r.h5")))
rng = np.random.default_rng(11)
host = torch.from_numpy(rng.integers(0, 256, size=(B, S, S, 3), dtype=np.uint8)).pin_memory()
def batches(nb):
    for _ in range(nb):
        yield host
for name, model in (("random", bench.build_model(S)), ("trained", trained())):
    det = blazeFaceDetector(model=model, inputSize=S)
    for packed in (False, True):
        for r in det.detect_stream(batches(3), 100, packed=packed): pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in det.detect_stream(batches(20), 100, packed=packed): pass
        torch.cuda.synchronize()
        print(name, "packed" if packed else "padded", (time.perf_counter() - t0) / 20 * 1e3, "ms/step")
    d = host.cuda()
    o = det.detect_device(d)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): o = det.detect_device(d, out=o)
    torch.cuda.synchronize(); print(name, "device padded", (time.perf_counter() - t0) / 20 * 1e3)
