"""B200-native drop-in for the reference's `multi_scale_face_detector.py::MultiScaleFaceDetector` (:28-288).

Same constructor arguments, same methods and return types (`preprocess_image`, `detect_single_scale`,
`calculate_scale_weights`, `weighted_nms`, `detect_multi_scale`), plus the batched entry point the
reference lacks (`detect_multi_scale_batch`): all images x all scales go through ONE pyramid launch,
one backbone forward per scale (unchanged PyTorch), and ONE post-processing call (decode+filter,
batched NMS, rescale, cross-scale weighted NMS) with a single device->host read at the end.
No prints on the hot path, no per-image host synchronisation, no CPU fallback.
"""
import threading
import time

import numpy as np
import torch

from . import _lib, ops, yolo_head
from .general import check_img_size


def _resolve_device(device):
    if isinstance(device, torch.device):
        d = device
    else:
        s = str(device).strip().lower()
        if s in ("", "cuda"):
            d = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        elif s == "cpu":
            raise RuntimeError("MultiScaleFaceDetector (libfdms) has no CPU path; use the reference on CPU")
        elif s.isdigit():
            d = torch.device("cuda", int(s))
        else:
            d = torch.device(s)
    if d is None or d.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("MultiScaleFaceDetector (libfdms) needs a CUDA device; none is available")
    return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())


def _load_model(model_path, device):
    """multi_scale_face_detector.py:48 -> utils/preprocess_yolo_predict.py:381 load_yolo_model.  A built
    nn.Module is used as is; a checkpoint path needs the reference's own model classes importable."""
    if isinstance(model_path, torch.nn.Module):
        return model_path.to(device).eval()
    try:
        from models.experimental import attempt_load  # the user's reference checkout (unchanged backbone)
    except Exception as e:  # pragma: no cover - depends on the deployment
        raise RuntimeError("loading a checkpoint path needs the reference's `models` package on sys.path "
                           "(the backbone stays the reference's PyTorch code); pass a built nn.Module instead") from e
    return attempt_load(model_path, map_location=device).eval()


class MultiScaleFaceDetector:
    """Multi-scale test-time-augmentation face detector (multi_scale_face_detector.py:28)."""

    def __init__(self, model_path, device='', img_sizes=[640, 3840], conf_thres=0.5, iou_thres=0.5,
                 use_api_preprocess=False, *, cap=4096, out_cap=None, iou_mode=ops.IOU_TV_CUDA, half=True,
                 with_landmarks=False):
        self.device = _resolve_device(device)
        torch.cuda.set_device(self.device)
        self.model = _load_model(model_path, self.device)
        self.half = bool(half)  # msfd:49: half precision on CUDA
        if self.half:
            self.model.half()
        self.stride = int(self.model.stride.max())                                    # msfd:51
        self.img_sizes = [check_img_size(size, s=self.stride) for size in img_sizes]  # msfd:52
        m = self.model.module if hasattr(self.model, 'module') else self.model
        self.names = getattr(m, 'names', ['face'])
        self.conf_thres = conf_thres
        self.iou_thres = iou_thres
        self.use_api_preprocess = use_api_preprocess
        # libfdms specifics
        self.cap = int(cap)                    # candidate capacity per (image, scale); grown on overflow
        self.out_cap = int(out_cap) if out_cap else None
        self.iou_mode = iou_mode               # per-scale NMS flavour; the merge NMS is always TV_CPU (msfd:234)
        self.with_landmarks = bool(with_landmarks)
        self.head = yolo_head.patch_head(yolo_head.find_head(self.model))
        if not yolo_head.is_face_head(self.head):  # the reference's class is a face detector too: it reads det[:, :6] of a
            # 21-column row (msfd:151); other heads go through patch_head + general.non_max_suppression (compat level)
            raise NotImplementedError("MultiScaleFaceDetector runs the face head (na=3, no=21); got na=%d no=%d"
                                      % (self.head.na, self.head.no))
        self.strides, self.anchors_px = yolo_head.head_geometry(self.head)
        self._tls = threading.local()  # plan caches hold device buffers + a native ctx: one set per calling thread
        self._speculate_small = True   # see _plan: dense batches switch it off until a sparse one is seen again
        self.last_stats = {}

    # ------------------------------------------------------------------ geometry / preprocessing
    def _geoms(self, h, w, sizes=None):
        sizes = self.img_sizes if sizes is None else sizes
        if self.use_api_preprocess:  # ppyp:345-378: zero-padded square, letterbox(auto=False)
            s = max(h, w)
            return [_lib.letterbox_geometry(s, s, size, auto=False, stride=self.stride) for size in sizes], (s, s)
        return [_lib.letterbox_geometry(h, w, size, auto=True, stride=self.stride) for size in sizes], (h, w)

    @staticmethod
    def _read(img):
        if isinstance(img, (str, bytes)) or hasattr(img, "__fspath__"):
            import cv2
            img0 = cv2.imread(str(img))  # msfd:84 (BGR)
            if img0 is None:
                raise ValueError(f"Could not read image: {img}")  # msfd:86
            return img0
        return img

    def _pyramid(self, batch_dev, sizes=None, ctx=None):
        """(B, H, W, 3) uint8 BGR on the device -> one (B, 3, Hs, Ws) model input per scale, one launch."""
        _, h, w, _ = batch_dev.shape
        geoms, virt = self._geoms(h, w, sizes)
        dtype = torch.float16 if self.half else torch.float32
        # cv2 decodes BGR; both branches of the reference feed RGB to the model (msfd:96 / ppyp:363)
        outs = ops.pyramid(batch_dev, geoms, out_dtype=dtype, layout="chw", flip_rb=True, virt_hw=virt, ctx=ctx)
        return outs, geoms

    def preprocess_image(self, img_path, img_size):
        """msfd:69.  Returns (img_tensor (1,3,H,W) on the device, img0 BGR ndarray, img0.shape)."""
        img0 = self._read(img_path)
        dev = torch.from_numpy(np.ascontiguousarray(img0)).to(self.device)[None]
        outs, _ = self._pyramid(dev, [img_size])
        return outs[0], img0, img0.shape

    # ------------------------------------------------------------------ model
    def _raw_heads(self, x):
        """Backbone forward (unchanged PyTorch) returning the head's raw NCHW conv maps."""
        yolo_head.set_fused(self.head, True)
        try:
            with torch.no_grad():
                raw = self.model(x, augment=False) if self._accepts_augment() else self.model(x)
        finally:
            yolo_head.set_fused(self.head, False)
        return list(raw)

    def _accepts_augment(self):
        a = getattr(self, "_aug", None)
        if a is None:
            import inspect
            try:
                a = "augment" in inspect.signature(self.model.forward).parameters
            except (TypeError, ValueError):
                a = False
            self._aug = a
        return a

    # ------------------------------------------------------------------ fused batch path
    def _plan(self, scales_levels, in_h, orig_hw, lane=None):
        plans = self._thread_plans() if lane is None else lane.plans
        bs = scales_levels[0][0].shape[0]
        # a plan owns device buffers sized by the capacities and a native context that is ordered by one stream
        stream_id = 0 if lane is not None else torch.cuda.current_stream(self.device).cuda_stream
        ragged = isinstance(orig_hw, torch.Tensor)
        key = (bs, tuple(tuple(l.shape[2:]) for lv in scales_levels for l in lv), tuple(in_h), "ragged" if ragged else tuple(orig_hw),
               self.cap, self.out_cap or self.cap, float(self.conf_thres), float(self.iou_thres), self.iou_mode,
               self.with_landmarks, stream_id)
        plan = plans.get(key)
        if plan is None:
            out_cap = self.out_cap or self.cap
            plan = ops.PostPlan(scales_levels, self.strides, self.anchors_px, in_h, orig_hw, self.conf_thres,
                                self.iou_thres, self.cap, out_cap, iou_mode=self.iou_mode,
                                with_kpts=self.with_landmarks, ctx=None if lane is None else lane.ctx)
            if len(plans) > 8:
                plans.clear()
            plans[key] = plan
        else:
            plan.rebind(scales_levels)
            if ragged:  # same shapes, other images: point the plan at this batch's sizes
                plan._orig_hw = orig_hw
                plan.desc.orig_hw_dev = orig_hw.data_ptr()
        # speculate that no segment / image exceeds the single-CTA kernels' limit while the recent batches did not: the
        # large-segment launches are then skipped and the exact counts read back with the results verify the guess
        plan.small_only = self._speculate_small
        return plan

    def _stream_lanes(self, depth):
        """The lanes of detect_stream are kept per calling thread: streams, native contexts (device scratch), plans and
        pinned result buffers are created once, not per call -- cudaMalloc / cudaHostAlloc / cudaFree synchronise the
        device and would stall the pipeline at the start and the end of every call."""
        lanes = getattr(self._tls, "lanes", None)
        if lanes is None:
            lanes = self._tls.lanes = []
        while len(lanes) < depth:
            lanes.append(self._Lane(self.device))
        return lanes[:depth]

    def _thread_plans(self):
        plans = getattr(self._tls, "plans", None)
        if plans is None:
            plans = self._tls.plans = {}
        return plans

    def detect_batch_device(self, batch_dev, lane=None):
        """Enqueue the whole multi-scale path for a (B, H, W, 3) uint8 BGR device batch on the current stream.
        Returns the plan whose `.out` (dets, n_out) and `.cands` (count, n_keep) stay on the device -- no
        synchronisation here.  `lane` (see detect_stream) supplies its own native context and plan cache."""
        _, h, w, _ = batch_dev.shape
        xs, geoms = self._pyramid(batch_dev, ctx=None if lane is None else lane.ctx)
        scales_levels = [self._raw_heads(x) for x in xs]
        plan = self._plan(scales_levels, [g.dst_h for g in geoms], (h, w), lane)
        plan.run()
        return plan

    @staticmethod
    def _meta(plan):
        """Device vector [n_out..., max count, min n_keep, max merged_in, sum count, sum n_keep] (one D2H)."""
        o, c = plan.out, plan.cands
        return torch.cat((o.n_out, torch.stack((c.count.max(), c.n_keep.min(), o.merged_in.max(), c.count.sum(),
                                                c.n_keep.sum())).to(torch.int32)))

    class _Lane:
        """Resources of one in-flight batch of detect_stream: stream, native context, plan cache, pinned results."""

        def __init__(self, device):
            self.stream = torch.cuda.Stream(device)
            self.ctx = _lib.Ctx(device.index)
            self.plans = {}
            self.img_dev = self.h_dets = self.h_meta = self.h_kpts = None
            self.done = torch.cuda.Event()
            self.plan = None
            self.src = None
            self.small_only = False

    def detect_stream(self, batches, depth=2):
        """Pipelined detection over an iterable of (B, H, W, 3) uint8 BGR host batches (pinned memory overlaps best):
        the H2D copy and kernels of batch i+1 are enqueued while batch i is in flight and batch i-1 is being read
        back, on `depth` streams with one native context each.  Yields, in order, the same per-batch result as
        detect_multi_scale_batch.  A batch that exceeds a capacity is transparently redone synchronously (the lanes'
        plans are rebuilt with the grown capacities afterwards).  With `with_landmarks`, `last_landmarks` / `last_stats`
        always describe the batch that was yielded last."""
        lanes = self._stream_lanes(max(1, int(depth)))
        for lane in lanes:
            lane.plan = lane.src = None
        pending = []

        def finish(lane):
            lane.done.synchronize()
            meta = lane.h_meta.numpy()
            n_out, head = meta[:-5], meta[-5:]
            small = ops.PostPlan.small_ok(head[0], head[2])
            missed = lane.small_only and not small
            self._speculate_small = small
            if head[0] > lane.plan.cands.cap or head[2] > lane.plan.out.out_cap or head[1] == -1 or missed:
                return self.detect_multi_scale_batch(lane.src)  # rare: grow capacities / full path, synchronously
            kmax = int(n_out.max()) if n_out.size else 0
            dets = lane.h_dets[:, :kmax].numpy().astype(np.float64)
            if lane.h_kpts is not None:  # with_landmarks: the landmarks of the batch being handed out
                kp = lane.h_kpts[:, :kmax].numpy().copy()
                self.last_landmarks = [kp[b, :n_out[b]] for b in range(len(n_out))]
            else:
                self.last_landmarks = None
            self.last_stats = {"candidates": int(head[3]), "kept_per_scale": int(head[4]), "final": int(n_out.sum())}
            return [dets[b, :n_out[b]] if n_out[b] else np.array([]) for b in range(len(n_out))]

        for i, batch in enumerate(batches):
            lane = lanes[i % len(lanes)]
            if lane.plan is not None and any(l is lane for l in pending):
                pending.remove(lane)
                yield finish(lane)
            t = batch if isinstance(batch, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(batch))
            if t.dim() != 4 or t.shape[3] != 3 or t.dtype != torch.uint8:
                raise ValueError("detect_stream: expected (B, H, W, 3) uint8 batches")
            with torch.cuda.stream(lane.stream):
                if t.is_cuda:  # already on the device (datasets.read_batch_device): no copy; the producer has synchronised
                    src_dev = t
                else:
                    if lane.img_dev is None or lane.img_dev.shape != t.shape:
                        lane.img_dev = torch.empty(t.shape, dtype=torch.uint8, device=self.device)
                    lane.img_dev.copy_(t, non_blocking=True)
                    src_dev = lane.img_dev
                plan = self.detect_batch_device(src_dev, lane)
                o = plan.out
                if lane.h_dets is None or lane.h_dets.shape != o.dets.shape or lane.h_meta.numel() != o.n_out.numel() + 5:
                    lane.h_dets = torch.empty(o.dets.shape, dtype=torch.float32).pin_memory()
                    lane.h_meta = torch.empty((o.n_out.numel() + 5,), dtype=torch.int32).pin_memory()
                lane.h_dets.copy_(o.dets, non_blocking=True)
                if o.kpts is not None:
                    if lane.h_kpts is None or lane.h_kpts.shape != o.kpts.shape:
                        lane.h_kpts = torch.empty(o.kpts.shape, dtype=torch.float32).pin_memory()
                    lane.h_kpts.copy_(o.kpts, non_blocking=True)
                else:
                    lane.h_kpts = None
                lane.h_meta.copy_(self._meta(plan), non_blocking=True)
                lane.done.record(lane.stream)
            lane.plan, lane.src, lane.small_only = plan, t, plan.small_only
            pending.append(lane)
        for lane in pending:
            yield finish(lane)

    def _collect(self, plan, device_out=False):
        """One device->host read of the results of `plan`; raises if a capacity was exceeded.  `device_out`: only the
        counts are read, the rows stay where they are."""
        o, c = plan.out, plan.cands
        both = self._meta(plan).cpu().numpy()
        n_out, head = both[:-5], both[-5:]  # one read, synchronises
        small = ops.PostPlan.small_ok(head[0], head[2])
        missed = plan.small_only and not small
        self._speculate_small = small
        if head[0] > c.cap or head[2] > o.out_cap or missed:
            return None, head
        if head[1] == -1:
            raise _lib.FdmsError("a segment exceeded the NMS suppression-mask budget")
        kmax = int(n_out.max()) if n_out.size else 0
        if device_out:
            return (o.dets[:, :kmax], o.n_out, o.kpts[:, :kmax] if o.kpts is not None else None), head
        dets = o.dets[:, :kmax].cpu().numpy().astype(np.float64) if kmax else np.zeros((len(n_out), 0, 7))
        kpts = o.kpts[:, :kmax].cpu().numpy() if (o.kpts is not None and kmax) else None
        return (dets, n_out, kpts), head

    def detect_multi_scale_batch(self, images, device_out=False):
        """images: (B, H, W, 3) uint8 BGR ndarray / tensor (pinned host or device), or a list of same-shape
        images / paths.  Returns a list of (k, 7) float64 arrays [x1,y1,x2,y2,conf,cls,scale_idx] per image, rows
        in weighted-score order (msfd:242-288); images without detections give np.array([]) (msfd:285).
        `device_out`: the detections stay on the GPU -- returns (dets (B, kmax, 7) fp32, n_out (B,) int32) tensors
        that are valid until the calling thread's next call (only the counts cross PCIe)."""
        if isinstance(images, (list, tuple)):
            from .datasets import read_batch, read_images
            if device_out:
                images = read_batch(images)
            else:
                decoded = read_images(images)  # parallel host decode, once per image
                if len({im.shape for im in decoded}) > 1:
                    return self._detect_ragged(decoded)  # mixed sizes: bucketed by letterboxed shape
                images = read_batch(decoded)  # one pinned batch
        t = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
        if t.dim() != 4 or t.shape[3] != 3 or t.dtype != torch.uint8:
            raise ValueError("detect_multi_scale_batch: expected (B, H, W, 3) uint8 images")
        batch_dev = t.to(self.device, non_blocking=True)
        t0 = time.perf_counter()
        while True:
            plan = self.detect_batch_device(batch_dev)
            res, head = self._collect(plan, device_out)
            if res is not None:
                break
            # capacity exceeded: grow to what this batch needs and redo (counts are exact even on overflow)
            if head[0] > self.cap:
                self.cap = max(self.cap * 2, int(head[0]))
            if head[2] > (self.out_cap or self.cap):
                self.out_cap = max(2 * (self.out_cap or self.cap), int(head[2]))
            self._thread_plans().clear()
            for lane in getattr(self._tls, "lanes", None) or []:  # the streaming lanes of this thread size theirs anew
                lane.plans.clear()  # (a batch in flight keeps its plan and its pinned buffers until it is read)
        dets, n_out, kpts = res
        if device_out:
            self.last_stats = {"seconds": time.perf_counter() - t0, "candidates": int(head[3]), "kept_per_scale": int(head[4])}
            self.last_landmarks = kpts
            return dets, n_out
        self.last_stats = {"seconds": time.perf_counter() - t0, "candidates": int(head[3]),
                           "kept_per_scale": int(head[4]), "final": int(n_out.sum())}
        self.last_landmarks = [kpts[b, :n_out[b]] for b in range(len(n_out))] if kpts is not None else None
        return [dets[b, :n_out[b]] if n_out[b] else np.array([]) for b in range(len(n_out))]

    def _detect_ragged(self, images):
        """Images of DIFFERENT sizes in one call (the reference takes arbitrary images one at a time,
        multi_scale_face_detector.py:242-288; autoShape batches mixed sizes, models/common.py:618-636).  They are
        bucketed by the shapes their letterboxed inputs have at every scale, so that the backbone and the whole
        post-processing chain still run once per bucket: inside a bucket, images of the same source size share a pyramid
        launch that writes straight into the bucket's batch tensors, and the merge takes every image's own original size
        (fdms_post_desc.orig_hw_dev).  Returns the per-image results in input order."""
        n = len(images)
        geoms = [self._geoms(im.shape[0], im.shape[1]) for im in images]
        buckets = {}
        for i, (g, _) in enumerate(geoms):
            buckets.setdefault(tuple((q.dst_h, q.dst_w) for q in g), []).append(i)
        results = [None] * n
        landmarks = [None] * n if self.with_landmarks else None
        stats = {"seconds": 0.0, "candidates": 0, "kept_per_scale": 0, "final": 0, "buckets": len(buckets)}
        dtype = torch.float16 if self.half else torch.float32
        t0 = time.perf_counter()
        for shapes, idx in buckets.items():
            idx = sorted(idx, key=lambda i: images[i].shape)  # same source size -> neighbours -> one launch
            nb = len(idx)
            outs = [torch.empty((nb, 3, h, w), dtype=dtype, device=self.device) for h, w in shapes]
            a = 0
            while a < nb:
                b = a
                while b < nb and images[idx[b]].shape == images[idx[a]].shape:
                    b += 1
                src = torch.from_numpy(np.stack([np.ascontiguousarray(images[i]) for i in idx[a:b]])).to(self.device, non_blocking=True)
                g, virt = geoms[idx[a]]
                ops.pyramid(src, g, out_dtype=dtype, layout="chw", flip_rb=True, virt_hw=virt, outs=[o[a:b] for o in outs])
                a = b
            hw = torch.tensor([[images[i].shape[0], images[i].shape[1]] for i in idx], dtype=torch.int32).to(self.device)
            scales_levels = [self._raw_heads(x) for x in outs]
            while True:
                plan = self._plan(scales_levels, [h for h, _ in shapes], hw)
                plan.run()
                res, head = self._collect(plan)
                if res is not None:
                    break
                if head[0] > self.cap:
                    self.cap = max(self.cap * 2, int(head[0]))
                if head[2] > (self.out_cap or self.cap):
                    self.out_cap = max(2 * (self.out_cap or self.cap), int(head[2]))
                self._thread_plans().clear()
            dets, n_out, kpts = res
            for j, i in enumerate(idx):
                results[i] = dets[j, :n_out[j]] if n_out[j] else np.array([])
                if landmarks is not None:
                    landmarks[i] = kpts[j, :n_out[j]]
            stats["candidates"] += int(head[3])
            stats["kept_per_scale"] += int(head[4])
            stats["final"] += int(n_out.sum())
        stats["seconds"] = time.perf_counter() - t0
        self.last_stats, self.last_landmarks = stats, landmarks
        return results

    # ------------------------------------------------------------------ reference-shaped methods
    @torch.no_grad()
    def detect_single_scale(self, img_path, img_size):
        """msfd:110.  Returns (detections (n,7) float64 [x1,y1,x2,y2,conf,cls,scale_idx], img0_shape, seconds)."""
        img0 = self._read(img_path)
        img0_shape = img0.shape
        dev = torch.from_numpy(np.ascontiguousarray(img0)).to(self.device)[None]
        xs, geoms = self._pyramid(dev, [img_size])
        torch.cuda.synchronize(self.device)  # time_synchronized (msfd:126)
        t1 = time.time()
        levels = self._raw_heads(xs[0])
        cap = self.cap
        while True:
            c = ops.decode_filter([levels], self.strides, self.anchors_px, self.conf_thres, cap)
            ops.nms_segments(c.cand, c.count, cap, self.iou_thres, self.iou_mode, tie=c.row, keep=c.keep, n_keep=c.n_keep)
            n_cand, nk = int(c.count[0]), int(c.n_keep[0])
            if n_cand <= cap:
                break
            cap = n_cand
        t2 = time.time()
        if nk < 0:
            raise _lib.FdmsError("a segment exceeded the NMS suppression-mask budget")
        if nk == 0:
            return np.zeros((0, 7)), img0_shape, t2 - t1  # msfd:164
        det = ops.gather_rows(c.cand, c.keep, c.n_keep, 6, nk)[0]
        ops.scale_coords_api_((geoms[0].dst_h, geoms[0].dst_w), det, img0_shape, do_round=True)  # msfd:144
        det_numpy = det.cpu().numpy().astype(np.float64)
        scale_idx = np.ones((det_numpy.shape[0], 1)) * self.img_sizes.index(img_size)  # msfd:156
        return np.hstack((det_numpy, scale_idx)), img0_shape, t2 - t1

    def calculate_scale_weights(self, detections):
        """msfd:168.  (n,7) -> (n,) float64 weights (1.0 / 1.1 / 1.2 by face area and scale index)."""
        detections = np.asarray(detections, np.float64)
        if len(detections) == 0:
            return np.array([])
        d = torch.from_numpy(np.ascontiguousarray(detections.reshape(-1, 7))).to(self.device)
        _, w = ops.weighted_rows(d, len(self.img_sizes), want_weights=True)
        return w.cpu().numpy()

    def weighted_nms(self, detections, iou_thresh=None):
        """msfd:203.  Size-aware score re-weighting then NMS; returns detections[keep] (original confidences,
        weighted-score order)."""
        detections = np.asarray(detections)
        if len(detections) == 0:
            return detections
        if iou_thresh is None:
            iou_thresh = self.iou_thres
        d = torch.from_numpy(np.ascontiguousarray(detections, dtype=np.float64)).to(self.device)
        n = d.shape[0]
        rows = ops.weighted_rows(d, len(self.img_sizes))
        count = torch.full((1,), n, dtype=torch.int32, device=self.device)
        keep, n_keep = ops.nms_segments(rows, count, n, float(iou_thresh), ops.IOU_TV_CPU, max_nms=0)
        nk = int(n_keep[0])
        if nk < 0:
            raise _lib.FdmsError("weighted_nms: the NMS suppression-mask budget was exceeded")
        return detections[keep[0, :nk].cpu().numpy()]

    def detect_multi_scale(self, img_path):
        """msfd:242.  Returns (final_detections (k,7) float64 or np.array([]), img0_shape)."""
        img0 = self._read(img_path)
        out = self.detect_multi_scale_batch(img0[None])
        return out[0], img0.shape

    def detect_files(self, paths, batch_size=32, decode_threads=None, depth=2, device_decode=False):
        """Generator over image files: yields `(path, detections (k,7) float64 or np.array([]))` in input order.  The
        production driver of the reference builds a detector per item and decodes every file once per scale
        (yolov7_face_multi_scale_dataframe_predict.py:607-700); here batch i+1 is decoded on host threads
        (datasets.read_images, cv2.imread: the reference's pixels) while batch i runs on the GPU.  Runs of batches whose
        images share one size go through detect_stream (pinned buffers recycled, H2D / kernels / D2H pipelined over
        `depth` streams); a batch of mixed sizes is bucketed by letterboxed shape (_detect_ragged).
        `device_decode=True`: JPEG files are only Huffman-decoded on the host threads; IDCT, chroma upsampling and colour
        conversion run on the GPU (datasets.read_batch_device, bit-identical to cv2.imread) and the batch never exists in
        host memory as pixels.  Files the device path does not take are decoded with cv2.imread into their slot."""
        import concurrent.futures
        from .datasets import read_batch, read_images
        paths = list(paths)
        chunks = [paths[i:i + batch_size] for i in range(0, len(paths), batch_size)]
        if not chunks:
            return
        # pinned / device batch buffers, recycled per (slot, batch length); depth + 2 slots are never in flight together.
        # They and the loader thread (whose libfdms context owns the pinned staging of the device JPEG path) outlive the
        # call: a new thread per call re-allocated ~400 MB of pinned and device memory every time.
        if not hasattr(self, "_file_bufs"):
            self._file_bufs, self._file_dev_bufs, self._file_active = {}, {}, False
            self._file_loader = concurrent.futures.ThreadPoolExecutor(1, thread_name_prefix="fdms-loader")
        shared = not self._file_active  # a second generator running on the same detector gets buffers of its own
        if shared:
            self._file_active = True
            bufs, dev_bufs, ex = self._file_bufs, self._file_dev_bufs, self._file_loader
        else:
            bufs, dev_bufs, ex = {}, {}, concurrent.futures.ThreadPoolExecutor(1)

        def load(ci):
            if device_decode:
                from .datasets import read_batch_device
                key = ci % (depth + 2)
                try:
                    dev_bufs[key], _ = read_batch_device(chunks[ci], device=self.device, threads=decode_threads or 0,
                                                         out=dev_bufs.get(key))
                    return dev_bufs[key]
                except ValueError as e:
                    if "the batch is" not in str(e):
                        raise  # unreadable file: same error as the host path
                    # mixed sizes: the host path buckets them
            imgs = read_images(chunks[ci], threads=decode_threads)
            if len({im.shape for im in imgs}) > 1:
                return imgs  # ragged
            key = (ci % (depth + 2), len(imgs))
            bufs[key] = read_batch(imgs, out=bufs.get(key) if bufs.get(key) is not None and
                                   tuple(bufs[key].shape[1:]) == imgs[0].shape else None)
            return bufs[key]

        try:
            state = {"next_ci": 0, "future": ex.submit(load, 0), "held": None}

            def take():
                """The next decoded chunk (and start decoding the one after it)."""
                if state["held"] is not None:
                    cur, state["held"] = state["held"], None
                    return cur
                if state["next_ci"] >= len(chunks):
                    return None
                ci = state["next_ci"]
                cur = state["future"].result()
                state["next_ci"] = ci + 1
                if ci + 1 < len(chunks):
                    state["future"] = ex.submit(load, ci + 1)
                return ci, cur

            def uniform_run(first):
                """`first` and the uniform batches that follow it; a ragged one is put back for the outer loop."""
                item = first
                while item is not None:
                    ci, cur = item
                    if isinstance(cur, list):
                        state["held"] = item
                        return
                    order.append(ci)
                    yield cur
                    item = take()

            while True:
                item = take()
                if item is None:
                    break
                ci, cur = item
                if isinstance(cur, list):
                    for pth, d in zip(chunks[ci], self._detect_ragged(cur)):
                        yield pth, d
                    continue
                order = []
                for k, dets in enumerate(self.detect_stream(uniform_run(item), depth=depth)):
                    for pth, d in zip(chunks[order[k]], dets):
                        yield pth, d
        finally:
            if shared:
                self._file_active = False
            else:
                ex.shutdown(wait=False)

    def get_detection_statistics(self, detections):
        """Subset of the reference's statistics helper used by detect_multi_scale's log line (msfd:281-282)."""
        detections = np.asarray(detections)
        if detections.size == 0:
            return {"total_faces": 0, "avg_confidence": 0.0}
        return {"total_faces": int(detections.shape[0]), "avg_confidence": float(detections[:, 4].mean())}


MultiScaleFaceDetectorB200 = MultiScaleFaceDetector
