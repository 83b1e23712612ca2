"""Detector facade with the surface of the reference's BlazePoser/blazeFaceDetectorH5.py.

``blazeFaceDetector`` (:80-357) and ``Results`` (:359-364) keep their names, constructor arguments,
attributes (``anchors``, ``inputHeight``, ``inputWidth``, ``sigmoidScoreThreshold``) and methods
(``detectFaces``, ``inference``, ``filterDetections``, ``extractDetections``,
``filterWithNonMaxSupression``, ``generateAnchors``).  Every numeric step runs in CUDA through
libhpose; the extra ``detectFacesBatch`` is the batched form the B200 path is built for.

``detectFaces(frame)`` is the latency path of the reference's webcam loop (:392-444): ONE host->device copy of the
uint8 frame (any size: the bicubic resize of :255 runs on the device), ONE CUDA-graph launch of the whole graph
(``hp_detect_frames``), ONE device->host copy of the packed result.  ``detectFacesStepwise`` walks the reference's five
methods one after the other (same results; the parity tests compare the two).  ``EMAFilter`` (:16-35) smooths a pose angle
over frames; drawing and the cv2 window loop (:37-77,128-219,366-449) are UI and stay out of scope.
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


FACE_DTYPE = np.dtype([("box", "<f8", (4,)), ("keypoints", "<f8", (KEY_POINT_SIZE, 2)), ("score", "<f4"), ("pose", "<f4", (3,)),
                       ("anchor", "<i4"), ("frame", "<i4")])          # hp_face (include/hpose.h), 152 bytes
assert FACE_DTYPE.itemsize == 152


class EMAFilter:
    """Exponential moving average of one scalar (reference :16-35): the first measurement initialises the state, then
    ``state = alpha * measurement + (1 - alpha) * state``."""

    def __init__(self, alpha: float, initial_value: float = 0.0):
        assert 0.0 < alpha <= 1.0, "alpha must be in (0,1]"
        self.alpha = alpha
        self.state = initial_value
        self.initialized = False

    def update(self, measurement: float) -> float:
        if self.initialized:
            self.state = self.alpha * measurement + (1.0 - self.alpha) * self.state
        else:
            self.state, self.initialized = measurement, True
        return self.state


def unpack_results(raw_u8, B, faces_off):
    """Packed hp_detect_frames result (host bytes) -> list of ``Results`` (copies: the buffer is reused by the next call)."""
    hdr = raw_u8[:(_lib.HP_RESULT_HEADER_INTS + B) * 4].view(np.int32)
    written = int(hdr[1])
    counts = hdr[_lib.HP_RESULT_HEADER_INTS:_lib.HP_RESULT_HEADER_INTS + B]
    faces = raw_u8[faces_off:faces_off + written * FACE_DTYPE.itemsize].view(FACE_DTYPE)
    out, o = [], 0
    for c in counts:
        f = faces[o:min(o + int(c), written)]
        out.append(Results(f["box"].copy(), f["keypoints"].copy(), f["score"].copy(), f["pose"].copy()))
        o += int(c)
    return out


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
        self._slots = {}
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

    # ------------------------------------------------------------------ single image
    def detectFaces(self, image):
        """One BGR uint8 frame of any size -> ``Results`` (reference :109-126) through the latency path: one pinned
        host->device copy, one CUDA-graph launch (``hp_detect_frames``), one device->host copy of the packed result."""
        results = self._detect_packed(np.ascontiguousarray(image)[None], graph=True)[0]
        self.updateFps()
        return results

    def detectFacesStepwise(self, image):
        """The reference's own sequence of calls (:109-126), every step a CUDA call with its own host round trip."""
        input_tensor = self.prepareInputForInference(image)
        loc_concat, cls_concat, pose_front, pose_back = self.inference(input_tensor)
        scores, good = self.filterDetections(cls_concat)
        boxes, keypoints = self.extractDetections(loc_concat, good)
        results = self.filterWithNonMaxSupression(boxes, keypoints, scores, good, pose_front, pose_back)
        self.updateFps()
        return results

    def prepareInputForInference(self, image):
        """BGR uint8 HxWx3 of any size -> (1,inputHeight,inputWidth,3) float32 in [-1,1] (reference :247-269: BGR->RGB,
        /255, bicubic resize, (x-0.5)/0.5), computed on the GPU (hp_preprocess_resize_u8)."""
        import torch
        image = np.ascontiguousarray(image)
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise ValueError("expected a HxWx3 uint8 BGR image")
        self.img_height, self.img_width, self.img_channels = image.shape
        x = self._preprocess_device(torch.from_numpy(image[None]).to(self.ctx.torch_device))
        return x.cpu().numpy()

    def _preprocess_device(self, u8):
        import torch
        B, H, W, _ = u8.shape
        x = torch.empty((B, self.inputHeight, self.inputWidth, 3), dtype=torch.float32, device=u8.device)
        _lib.check(_lib.lib().hp_preprocess_resize_u8(self.ctx.handle, u8.data_ptr(), B, H, W, self.inputHeight, self.inputWidth,
                                                      x.data_ptr(), self.ctx.stream_ptr()))
        return x

    # ------------------------------------------------------------------ packed single-call path (hp_detect_frames)
    def _latency_slot(self, B, H, W, max_faces):
        """Persistent buffers of one (batch, frame size): pinned host frame + result, device frame + result, a private
        non-default stream (graph capture needs one).  Pointers stay fixed, so the captured graph stays valid."""
        import torch
        key = (B, H, W, max_faces)
        slot = self._slots.get(key)
        if slot is None:
            if len(self._slots) >= 4:
                self._slots.clear()
            dev = self.ctx.torch_device
            L = _lib.lib()
            nbytes = int(L.hp_detect_result_bytes(B, B * max_faces))
            slot = {"h_in": torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory(),
                    "d_in": torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev),
                    "d_out": torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev),
                    "h_out": torch.empty((nbytes + 7) // 8, dtype=torch.int64).pin_memory(),
                    "faces_off": int(L.hp_detect_result_faces_offset(B)), "nbytes": nbytes,
                    "stream": torch.cuda.Stream(dev)}
            self._slots[key] = slot
        return slot

    def _detect_packed(self, frames, max_faces=MAX_FACE_NUM, graph=False):
        """frames: (B,H,W,3) uint8 BGR host array -> list of ``Results``.  One H2D, one launch sequence (or graph), one D2H
        (two for large batches: the header first, then exactly sum(count) face records)."""
        import torch
        if frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] != 3:
            raise ValueError("expected (B,H,W,3) uint8 BGR frames")
        B, H, W, _ = frames.shape
        self.img_height, self.img_width, self.img_channels = H, W, 3
        slot = self._latency_slot(B, H, W, max_faces)
        m = self.interpreter.to_device(self.ctx)
        st = slot["stream"]
        slot["h_in"].numpy()[...] = frames
        with torch.cuda.stream(st):
            slot["d_in"].copy_(slot["h_in"], non_blocking=True)
            _lib.check(_lib.lib().hp_detect_frames(
                self.ctx.handle, m.head16.head_handle, m.head8.head_handle, slot["d_in"].data_ptr(), B, H, W, self.inputHeight,
                self.inputWidth, float(np.float32(self.sigmoidScoreThreshold)), float(np.float32(self.iouThreshold)), int(max_faces),
                slot["d_out"].data_ptr(), slot["nbytes"], _lib.HP_DETECT_GRAPH if graph else 0, C.c_void_p(st.cuda_stream)))
            h_raw = slot["h_out"].numpy().view(np.uint8)
            if slot["nbytes"] <= (1 << 20):
                slot["h_out"].copy_(slot["d_out"], non_blocking=True)
                st.synchronize()
            else:
                n_hdr = (slot["faces_off"] + 7) // 8
                slot["h_out"][:n_hdr].copy_(slot["d_out"][:n_hdr], non_blocking=True)
                st.synchronize()
                written = int(h_raw[:16].view(np.int32)[1])
                n_all = n_hdr + (written * FACE_DTYPE.itemsize + 7) // 8
                if written:
                    slot["h_out"][n_hdr:n_all].copy_(slot["d_out"][n_hdr:n_all], non_blocking=True)
                    st.synchronize()
        return unpack_results(h_raw, B, slot["faces_off"])

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
        """images: (B,H,W,3) uint8 BGR frames of any size (host array or CUDA tensor) -> list of ``Results``; one fused
        device pass: preprocess (+ resize) -> backbone -> heads -> decode + NMS + pose lookup.  Host arrays take the packed
        single-call path (only sum(count) face records come back)."""
        if isinstance(images, np.ndarray):
            return self._detect_packed(np.ascontiguousarray(images), max_faces)
        out = self.detect_device(images, max_faces)
        cnt = out["count"].cpu().numpy()
        boxes, kps = out["boxes"].cpu().numpy(), out["keypoints"].cpu().numpy()
        scores, poses = out["scores"].cpu().numpy(), out["poses"].cpu().numpy()
        return [Results(boxes[i, :c].copy(), kps[i, :c].copy(), scores[i, :c].copy(), poses[i, :c].copy())
                for i, c in enumerate(cnt)]

    def detect_stream(self, host_batches, max_faces=MAX_FACE_NUM, keys=("count", "boxes", "keypoints", "scores", "poses"), chunks=1,
                      packed=False):
        """Pipelined serving loop: ``host_batches`` yields pinned (B,H,W,3) uint8 BGR host tensors; for every batch
        a dict of pinned HOST tensors is yielded, in order.  The host->device copy of batch i+1 and the device->host read of
        batch i-1 run on their own CUDA streams while batch i computes: PCIe time hides behind the kernels instead of
        adding to them.

        ``packed=False``: the dict holds ``keys`` in the padded [B, max_faces, ...] form of ``hp_unified_forward``.
        ``chunks`` > 1 splits every batch into that many slices which go through copy / compute / read-back one after the
        other (shorter pipeline fill and drain; the results are the same).
        ``packed=True``: ``hp_detect_frames``; the dict holds ``count`` (int32 [B]), ``faces`` (``FACE_DTYPE`` records of
        all frames, frame-major) and ``total``; only the header and sum(count) records cross PCIe (the padded form moves
        14.4 KB per frame whatever was found).

        Buffer lifetime: three result slots rotate and at most two batches are in flight, so a yielded dict stays valid
        until the generator has been advanced TWICE more (the slot being refilled is never the one just handed out).
        The slots (device input / result tensors, PINNED host result tensors, streams, events) belong to the detector and
        are reused by later ``detect_stream`` calls with the same ``packed`` / ``chunks``: pinning ~60 MB of host memory per
        slot costs ~6 ms, which a serving loop must not pay per call.  Results of an earlier call are therefore overwritten by
        the next one -- copy what has to live longer.  (A second generator started while one is still running gets private slots.)"""
        import torch
        dev = self.ctx.torch_device
        comp = torch.cuda.current_stream(dev)
        chunks = 1 if packed else max(1, int(chunks))
        NS = 3
        cache = self.__dict__.setdefault("_stream_slots", {})
        st = cache.get((bool(packed), chunks))
        if st is None or st["busy"]:
            st = {"busy": False,
                  "s_in": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                  "s_rec": torch.cuda.Stream(dev),      # second-phase record copies: must not queue behind the NEXT batch's read-back
                  "d_in": [[None] * chunks for _ in range(NS)],
                  "d_out": [[None] * chunks for _ in range(NS)],   # device result tensors, reused: slot b is idle once its read-back is done
                  "ev_read": [[None] * chunks for _ in range(NS)],
                  "h_out": [None] * NS,
                  "ev_in": [[torch.cuda.Event() for _ in range(chunks)] for _ in range(NS)],
                  "ev_comp": [[torch.cuda.Event() for _ in range(chunks)] for _ in range(NS)],
                  "used": [[False] * chunks for _ in range(NS)],
                  "ev_out": [torch.cuda.Event() for _ in range(NS)]}
            if (bool(packed), chunks) not in cache or not cache[(bool(packed), chunks)]["busy"]:
                cache[(bool(packed), chunks)] = st
        st["busy"] = True
        s_in, s_out, s_rec = st["s_in"], st["s_out"], st["s_rec"]
        d_in, d_out, ev_read, h_out = st["d_in"], st["d_out"], st["ev_read"], st["h_out"]
        ev_in, ev_comp, used, ev_out = st["ev_in"], st["ev_comp"], st["used"], st["ev_out"]
        meta = [None] * NS
        pending = []                                      # slots whose results are still on their way to the host
        L = _lib.lib()
        m = self.interpreter.to_device(self.ctx)

        def finish(q):
            ev_out[q].synchronize()
            if not packed:
                return h_out[q]
            B, faces_off, n_hdr = meta[q]
            raw = h_out[q].numpy().view(np.uint8)
            hdr = raw[:(_lib.HP_RESULT_HEADER_INTS + B) * 4].view(np.int32)
            written = int(hdr[1])
            if written:                                   # second phase: exactly the records that exist
                n_all = n_hdr + (written * FACE_DTYPE.itemsize + 7) // 8
                with torch.cuda.stream(s_rec):
                    h_out[q][n_hdr:n_all].copy_(d_out[q][0][n_hdr:n_all], non_blocking=True)
                    ev_read[q][0].record(s_rec)
                    ev_out[q].record(s_rec)
                ev_out[q].synchronize()
            return {"total": int(hdr[0]), "count": hdr[_lib.HP_RESULT_HEADER_INTS:_lib.HP_RESULT_HEADER_INTS + B],
                    "faces": raw[faces_off:faces_off + written * FACE_DTYPE.itemsize].view(FACE_DTYPE)}

        n = 0
        try:
            yield from self._detect_stream_loop(host_batches, max_faces, keys, chunks, packed, st, finish, pending, meta, comp, L, m, NS)
        finally:
            st["busy"] = False

    def _detect_stream_loop(self, host_batches, max_faces, keys, chunks, packed, st, finish, pending, meta, comp, L, m, NS):
        import torch
        dev = self.ctx.torch_device
        s_in, s_out = st["s_in"], st["s_out"]
        d_in, d_out, ev_read, h_out = st["d_in"], st["d_out"], st["ev_read"], st["h_out"]
        ev_in, ev_comp, used, ev_out = st["ev_in"], st["ev_comp"], st["used"], st["ev_out"]
        n = 0
        for hb in host_batches:
            b = n % NS
            if len(pending) == 2:                          # keep two batches in flight: hand the oldest out first
                yield finish(pending.pop(0))
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
                if packed:
                    nbytes = int(L.hp_detect_result_bytes(B, B * max_faces))
                    nwords = (nbytes + 7) // 8
                    if d_out[b][c] is None or d_out[b][c].numel() != nwords:
                        d_out[b][c] = torch.empty(nwords, dtype=torch.int64, device=dev)
                        h_out[b] = torch.empty(nwords, dtype=torch.int64).pin_memory()
                    _, H, W, _ = part.shape
                    _lib.check(L.hp_detect_frames(
                        self.ctx.handle, m.head16.head_handle, m.head8.head_handle, d_in[b][c].data_ptr(), B, H, W, self.inputHeight,
                        self.inputWidth, float(np.float32(self.sigmoidScoreThreshold)), float(np.float32(self.iouThreshold)),
                        int(max_faces), d_out[b][c].data_ptr(), nbytes, 0, self.ctx.stream_ptr()))
                    out = None
                else:
                    out = d_out[b][c] = self.detect_device(d_in[b][c], max_faces, out=d_out[b][c])
                ev_comp[b][c].record(comp)
                used[b][c] = True
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_comp[b][c])
                    if packed:
                        faces_off = int(L.hp_detect_result_faces_offset(B))
                        n_hdr = (faces_off + 7) // 8
                        meta[b] = (B, faces_off, n_hdr)
                        h_out[b][:n_hdr].copy_(d_out[b][c][:n_hdr], non_blocking=True)
                    else:
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
            yield finish(q)

    def detect_device(self, images, max_faces=MAX_FACE_NUM, float_input=False, out=None):
        import torch
        dev = self.ctx.torch_device
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images)).to(dev, non_blocking=True)
        if images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError("expected (B,H,W,3) images")
        x = images.float().contiguous() if float_input else self._preprocess_device(images.contiguous())
        B, H, W, _ = x.shape
        m = self.interpreter.to_device(self.ctx)
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
