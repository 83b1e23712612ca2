"""Detector facade with the surface of the reference's BlazePoser/blazeFaceDetectorH5.py.

``blazeFaceDetector`` (:80-357) and ``Results`` (:359-364) keep their names, constructor arguments,
attributes (``anchors``, ``inputHeight``, ``inputWidth``, ``sigmoidScoreThreshold``) and methods
(``detectFaces``, ``inference``, ``filterDetections``, ``extractDetections``,
``filterWithNonMaxSupression``, ``generateAnchors``).  Every numeric step runs in CUDA through
libhpose; the extra ``detectFacesBatch`` is the batched form the B200 path is built for.
The webcam loop, drawing and EMA smoothing of the reference (:16-77,128-219,366-449) are UI and out
of scope.
"""
import ctypes as C
import os
import time

import numpy as np

from . import _lib
from .blazeFaceUtils import SsdAnchorsCalculatorOptions, gen_anchors
from .device import default_context
from .unified import UnifiedModel

KEY_POINT_SIZE = 6
MAX_FACE_NUM = 100
INPUT_FRONT = 128
INPUT_BACK = 256
DEFAULT_MODEL = "UnifiedModels/reg1-stoqa9pt-reg2-hrchr82r-selected.h5"


class Results:
    def __init__(self, boxes, keypoints, scores, poses):
        self.boxes = boxes
        self.keypoints = keypoints
        self.scores = scores
        self.poses = poses


class blazeFaceDetector:
    def __init__(self, scoreThreshold=0.4, iouThreshold=0.3, modelPath=None, model=None, inputSize=INPUT_FRONT):
        self.scoreThreshold = scoreThreshold
        self.iouThreshold = iouThreshold
        self.sigmoidScoreThreshold = np.log(self.scoreThreshold / (1 - self.scoreThreshold))
        self.fps = 0
        self.timeLastPrediction = time.time()
        self.frameCounter = 0
        self._model_path = modelPath or DEFAULT_MODEL
        self._model = model
        self._input_size = int(inputSize)
        self.initializeModel()
        self.generateAnchors()

    # ------------------------------------------------------------------ set-up
    def initializeModel(self):
        self.interpreter = self._model if self._model is not None else UnifiedModel.load(self._model_path)
        self.ctx = default_context()
        self.interpreter.to_device(self.ctx)
        self.getModelInputDetails()

    def getModelInputDetails(self):
        self.inputHeight = self._input_size
        self.inputWidth = self._input_size
        self.channels = 3

    def generateAnchors(self):
        opts = SsdAnchorsCalculatorOptions(
            input_size_width=self.inputWidth, input_size_height=self.inputHeight, min_scale=0.1484375, max_scale=0.75,
            anchor_offset_x=0.5, anchor_offset_y=0.5, num_layers=4, feature_map_width=[], feature_map_height=[],
            strides=[8, 16, 16, 16], aspect_ratios=[1.0], reduce_boxes_in_lowest_layer=False,
            interpolated_scale_aspect_ratio=1.0, fixed_anchor_size=True)
        self.anchors = gen_anchors(opts)

    def updateFps(self):
        now = time.time()
        self.fps = int(1 / (now - self.timeLastPrediction + 0.0001))
        self.timeLastPrediction = now

    # ------------------------------------------------------------------ single image, reference flow
    def detectFaces(self, image):
        input_tensor = self.prepareInputForInference(image)
        loc_concat, cls_concat, pose_front, pose_back = self.inference(input_tensor)
        scores, good = self.filterDetections(cls_concat)
        boxes, keypoints = self.extractDetections(loc_concat, good)
        results = self.filterWithNonMaxSupression(boxes, keypoints, scores, good, pose_front, pose_back)
        self.updateFps()
        return results

    def prepareInputForInference(self, image):
        """BGR uint8 HxWx3 -> (1,H,W,3) float32 in [-1,1] on the GPU (hp_preprocess_u8).  Images must
        already have the network input size: the bicubic resize of the reference (:255) is a later row."""
        import torch
        image = np.ascontiguousarray(image)
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise ValueError("expected a HxWx3 uint8 BGR image")
        self.img_height, self.img_width, self.img_channels = image.shape
        if (self.img_height, self.img_width) != (self.inputHeight, self.inputWidth):
            raise ValueError(f"image is {self.img_width}x{self.img_height}; resize to {self.inputWidth}x"
                             f"{self.inputHeight} first (on-device bicubic resize is not built yet)")
        x = self._preprocess_device(torch.from_numpy(image[None]).to(self.ctx.torch_device))
        return x.cpu().numpy()

    def _preprocess_device(self, u8):
        import torch
        B, H, W, _ = u8.shape
        x = torch.empty((B, H, W, 3), dtype=torch.float32, device=u8.device)
        _lib.check(_lib.lib().hp_preprocess_u8(self.ctx.handle, u8.data_ptr(), B, H, W, x.data_ptr(),
                                               self.ctx.stream_ptr()))
        return x

    def inference(self, input_tensor):
        raw = self.interpreter(input_tensor)
        cls_front, cls_back = np.squeeze(raw[0]), np.squeeze(raw[1])
        loc_front, loc_back = np.squeeze(raw[2]), np.squeeze(raw[3])
        pose_front, pose_back = np.squeeze(raw[4]), np.squeeze(raw[5])
        axis = 0 if cls_front.ndim == 1 else 1
        return (np.concatenate((loc_front, loc_back), axis=axis), np.concatenate((cls_front, cls_back), axis=axis),
                pose_front, pose_back)

    def filterDetections(self, output1):
        import torch
        cls = torch.from_numpy(np.ascontiguousarray(output1, np.float32).reshape(1, -1)).to(self.ctx.torch_device)
        A = cls.shape[1]
        idx = torch.empty((1, A), dtype=torch.int32, device=cls.device)
        sc = torch.empty((1, A), dtype=torch.float32, device=cls.device)
        cnt = torch.zeros((1,), dtype=torch.int32, device=cls.device)
        _lib.check(_lib.lib().hp_filter_detections(self.ctx.handle, cls.data_ptr(), 1, A,
                                                   float(np.float32(self.sigmoidScoreThreshold)), idx.data_ptr(),
                                                   sc.data_ptr(), cnt.data_ptr(), self.ctx.stream_ptr()))
        n = int(cnt.item())
        return sc[0, :n].cpu().numpy(), idx[0, :n].cpu().numpy().astype(np.int64)

    def extractDetections(self, output0, goodDetectionsIndices):
        import torch
        n = int(len(goodDetectionsIndices))
        dev = self.ctx.torch_device
        loc = torch.from_numpy(np.ascontiguousarray(output0, np.float32)).to(dev)
        idx = torch.from_numpy(np.ascontiguousarray(goodDetectionsIndices, np.int32)).to(dev)
        boxes = torch.zeros((n, 4), dtype=torch.float64, device=dev)
        kps = torch.zeros((n, KEY_POINT_SIZE, 2), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().hp_extract_detections(self.ctx.handle, loc.data_ptr(), idx.data_ptr(), n, self.inputHeight,
                                                    self.inputWidth, boxes.data_ptr(), kps.data_ptr(),
                                                    self.ctx.stream_ptr()))
        return boxes.cpu().numpy(), kps.cpu().numpy()

    def filterWithNonMaxSupression(self, boxes, keypoints, scores, detection_indices, pose_front, pose_back):
        import torch
        dev = self.ctx.torch_device
        n = int(len(scores))
        sel = torch.zeros((MAX_FACE_NUM,), dtype=torch.int32, device=dev)
        cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
        b = torch.from_numpy(np.ascontiguousarray(boxes, np.float64).reshape(n, 4)).to(dev)
        s = torch.from_numpy(np.ascontiguousarray(scores, np.float32)).to(dev)
        _lib.check(_lib.lib().hp_nms(self.ctx.handle, b.data_ptr(), s.data_ptr(), n, float(np.float32(self.iouThreshold)),
                                     MAX_FACE_NUM, sel.data_ptr(), cnt.data_ptr(), self.ctx.stream_ptr()))
        selected = sel[:int(cnt.item())].cpu().numpy().astype(np.int64)
        boxes, keypoints, scores = np.asarray(boxes)[selected], np.asarray(keypoints)[selected], np.asarray(scores)[selected]
        if selected.size == 0:
            return Results(boxes, keypoints, scores, np.zeros((0, 3), dtype=np.float32))
        anchors = np.asarray(detection_indices)[selected]
        front = anchors < pose_front.shape[0] * pose_front.shape[1] * 2
        cell = np.where(front, anchors // 2, (anchors - pose_front.shape[0] * pose_front.shape[1] * 2) // 6)
        flat_front = pose_front.reshape(-1, pose_front.shape[-1])
        flat_back = pose_back.reshape(-1, pose_back.shape[-1])
        poses = np.where(front[:, None], flat_front[np.minimum(cell, len(flat_front) - 1)],
                         flat_back[np.minimum(cell, len(flat_back) - 1)])
        return Results(boxes, keypoints, scores, poses)

    # ------------------------------------------------------------------ batched B200 path
    def detectFacesBatch(self, images, max_faces=MAX_FACE_NUM):
        """images: (B,H,W,3) uint8 BGR (host array or CUDA tensor) -> list of ``Results``; one fused
        device pass: preprocess -> backbone -> heads -> decode + NMS + pose lookup."""
        out = self.detect_device(images, max_faces)
        cnt = out["count"].cpu().numpy()
        boxes, kps = out["boxes"].cpu().numpy(), out["keypoints"].cpu().numpy()
        scores, poses = out["scores"].cpu().numpy(), out["poses"].cpu().numpy()
        return [Results(boxes[i, :c].copy(), kps[i, :c].copy(), scores[i, :c].copy(), poses[i, :c].copy())
                for i, c in enumerate(cnt)]

    def detect_stream(self, host_batches, max_faces=MAX_FACE_NUM, keys=("count", "boxes", "keypoints", "scores", "poses"), chunks=1):
        """Pipelined serving loop: ``host_batches`` yields pinned (B,H,W,3) uint8 BGR host tensors; for every batch
        a dict of pinned HOST tensors (``keys``) is yielded, in order.  The host->device copy of batch i+1 and the
        device->host read of batch i-1 run on their own CUDA streams while batch i computes (two device input
        buffers, two sets of host result buffers): PCIe time hides behind the kernels instead of adding to them.
        ``chunks`` > 1 splits every batch into that many slices which go through copy / compute / read-back one after
        the other: the first kernels then wait for 1/chunks of the first copy only and the last read-back is 1/chunks of
        a batch (shorter pipeline fill and drain; the results are the same).
        A yielded dict is reused two batches later: consume (or copy) it before asking for the next-but-one."""
        import torch
        dev = self.ctx.torch_device
        comp = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        chunks = max(1, int(chunks))
        d_in = [[None] * chunks for _ in range(2)]
        d_out = [[None] * chunks for _ in range(2)]       # device result tensors, reused: slot b is idle once its read-back is done
        ev_read = [[None] * chunks for _ in range(2)]
        h_out = [None, None]
        ev_in = [[torch.cuda.Event() for _ in range(chunks)] for _ in range(2)]
        ev_comp = [[torch.cuda.Event() for _ in range(chunks)] for _ in range(2)]
        used = [[False] * chunks for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        pending = []                                      # slots whose results are still on their way to the host
        n = 0
        for hb in host_batches:
            b = n & 1
            if len(pending) == 2:                          # slot b is reused: hand its results out first
                q = pending.pop(0)
                ev_out[q].synchronize()
                yield h_out[q]
            B = hb.shape[0]
            per = -(-B // chunks)
            for c in range(chunks):
                c0, c1 = c * per, min(B, (c + 1) * per)
                if c0 >= c1:
                    break
                part = hb[c0:c1]
                if d_in[b][c] is None or d_in[b][c].shape != part.shape:
                    d_in[b][c] = torch.empty(part.shape, dtype=part.dtype, device=dev)
                with torch.cuda.stream(s_in):
                    if used[b][c]:
                        s_in.wait_event(ev_comp[b][c])     # the pass that read this input buffer is done
                    d_in[b][c].copy_(part, non_blocking=True)
                    ev_in[b][c].record(s_in)
                comp.wait_event(ev_in[b][c])
                if ev_read[b][c] is not None:
                    comp.wait_event(ev_read[b][c])         # the previous results of this slot have left the device
                out = d_out[b][c] = self.detect_device(d_in[b][c], max_faces, out=d_out[b][c])
                ev_comp[b][c].record(comp)
                used[b][c] = True
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_comp[b][c])
                    if h_out[b] is None or any(h_out[b][k].shape != (B,) + tuple(out[k].shape[1:]) for k in keys):
                        h_out[b] = {k: torch.empty((B,) + tuple(out[k].shape[1:]), dtype=out[k].dtype).pin_memory() for k in keys}
                    for k in keys:
                        h_out[b][k][c0:c1].copy_(out[k], non_blocking=True)
                    if ev_read[b][c] is None:
                        ev_read[b][c] = torch.cuda.Event()
                    ev_read[b][c].record(s_out)
            ev_out[b].record(s_out)
            pending.append(b)
            n += 1
        for q in pending:
            ev_out[q].synchronize()
            yield h_out[q]

    def detect_device(self, images, max_faces=MAX_FACE_NUM, float_input=False, out=None):
        import torch
        dev = self.ctx.torch_device
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images)).to(dev, non_blocking=True)
        if images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError("expected (B,H,W,3) images")
        x = images.float().contiguous() if float_input else self._preprocess_device(images.contiguous())
        B, H, W, _ = x.shape
        m = self.interpreter
        H16, W16, H8, W8 = -(-H // 8), -(-W // 8), -(-H // 16), -(-W // 16)
        if out is not None and out["count"].shape[0] == B and out["anchors"].shape[1] == max_faces and out["pose16"].shape[1:3] == (H16, W16):
            pass                                   # caller-owned result tensors of the right shape (the serving loop reuses them)
        else:
            out = {"count": torch.empty((B,), dtype=torch.int32, device=dev),
                   "anchors": torch.empty((B, max_faces), dtype=torch.int32, device=dev),
                   "boxes": torch.empty((B, max_faces, 4), dtype=torch.float64, device=dev),
                   "keypoints": torch.empty((B, max_faces, KEY_POINT_SIZE, 2), dtype=torch.float64, device=dev),
                   "scores": torch.empty((B, max_faces), dtype=torch.float32, device=dev),
                   "poses": torch.empty((B, max_faces, 3), dtype=torch.float32, device=dev),
                   "pose16": torch.empty((B, H16, W16, 3), dtype=torch.float32, device=dev),
                   "pose8": torch.empty((B, H8, W8, 3), dtype=torch.float32, device=dev)}
        _lib.check(_lib.lib().hp_unified_forward(
            self.ctx.handle, m.head16.head_handle, m.head8.head_handle, x.data_ptr(), B, H, W,
            float(np.float32(self.sigmoidScoreThreshold)), float(np.float32(self.iouThreshold)), int(max_faces),
            out["pose16"].data_ptr(), out["pose8"].data_ptr(), out["count"].data_ptr(), out["anchors"].data_ptr(),
            out["boxes"].data_ptr(), out["keypoints"].data_ptr(), out["scores"].data_ptr(), out["poses"].data_ptr(),
            self.ctx.stream_ptr()))
        return out
