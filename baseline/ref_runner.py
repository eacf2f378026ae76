"""Drives the UNMODIFIED reference over the bench workload (TEST / BASELINE INFRASTRUCTURE ONLY).

Only `tests/` and `bench.py`'s `cpu_baseline` / `--impl reference` legs import this module; nothing under
`face_detection_multi_scale_b200/` does.  The reference tree is found by `baseline/refshim.py`
(`/root/reference` in the build container, the byte-for-byte copy `baseline/_ref/` on the GPU box).

What runs is the reference's own code, function for function: `MultiScaleFaceDetector.detect_multi_scale`
(multi_scale_face_detector.py:242-288) -> `detect_single_scale` (:110-166) -> `preprocess_image` (:69-107: `letterbox`,
BGR->RGB, CHW, `/255`) -> `model(img)` -> `non_max_suppression` (utils/general.py:506) ->
`scale_coords_api_approach(...).round()` (utils/preprocess_yolo_predict.py:122) -> `weighted_nms` (:203-240).
Two things are supplied from outside, exactly as in `bench.py`'s own arm:

* the image comes from memory: the module-level name `cv2` of `multi_scale_face_detector` is bound to a proxy whose
  `imread("mem:<i>")` returns the i-th array of the batch (everything else forwards to OpenCV), so no file is decoded
  inside the timed region on either arm;
* the backbone (out of scope, unchanged PyTorch) is a stand-in: `model(img)` hands the pre-generated raw head maps
  `(1, 63, ny, nx)` of that image and scale to the reference's own `IKeypoint` head (models/yolo.py:215-313), whose
  convolutions are replaced by channel slices (`m`: channels 0..17, `m_kpt`: 18..62, `ia` / `im`: identity) -- the
  decode after them is the reference's code, untouched.

Modes: `as_written` = one process, batch 1, sequential scales, default torch threads,
OpenCV single-threaded (utils/general.py:31); `all_cores` = one spawned single-thread worker process per host core,
image-sharded, like the reference's own production driver (yolov7_face_multi_scale_dataframe_predict.py:1115,1147).
"""
import contextlib
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_DEVNULL = None


@contextlib.contextmanager
def _quiet():
    """The reference prints per image and per scale (msfd:258-283, ppyp:138-156): keep the work, drop the text."""
    global _DEVNULL
    if _DEVNULL is None:
        _DEVNULL = open(os.devnull, "w")
    with contextlib.redirect_stdout(_DEVNULL):
        yield


def available():
    from baseline import refshim
    return refshim.find_root() is not None


class _Cv2Mem:
    """`cv2` as seen by multi_scale_face_detector.py: `imread('mem:<key>')` serves an in-memory BGR array."""

    def __init__(self, cv2, store):
        self._cv2, self._store = cv2, store

    def imread(self, path, *a):
        if isinstance(path, str) and path.startswith("mem:"):
            return self._store[path]
        return self._cv2.imread(path, *a)

    def __getattr__(self, name):
        return getattr(self._cv2, name)


class RefHarness:
    """One reference `MultiScaleFaceDetector` around the stand-in backbone (see module docstring)."""

    def __init__(self, img_sizes, conf_thres, iou_thres, strides, anchors_px, device="cpu", model=None):
        from baseline import refshim
        if not refshim.install():
            raise RuntimeError("no reference tree (neither /root/reference nor baseline/_ref)")
        import torch
        warnings.filterwarnings("ignore", message="torch.meshgrid")  # models/yolo.py:311 (old call signature)
        with _quiet():
            import multi_scale_face_detector as msfd
            from models.yolo import IKeypoint
            from utils.general import check_img_size
        import cv2
        self.torch, self.msfd = torch, msfd
        self.store = {}
        msfd.cv2 = _Cv2Mem(cv2, self.store)
        self.device = torch.device(device)
        nl = len(strides)

        class _Slice(torch.nn.Module):
            def __init__(self, a, b):
                super().__init__()
                self.a, self.b = a, b

            def forward(self, x):
                return x[:, self.a:self.b]

        class _StandInBackbone(torch.nn.Module):
            """Delivers the raw head maps of the current (image, input size) to the reference's IKeypoint."""

            def __init__(s):
                super().__init__()
                with _quiet():
                    s.head = IKeypoint(nc=1, anchors=np.asarray(anchors_px, np.float32).reshape(nl, -1).tolist(), nkpt=5,
                                       ch=(63,) * nl)
                na, no_det = s.head.na, s.head.no_det
                for i in range(nl):
                    s.head.ia[i] = torch.nn.Identity()
                    s.head.im[i] = torch.nn.Identity()
                    s.head.m[i] = _Slice(0, na * no_det)
                    s.head.m_kpt[i] = _Slice(na * no_det, na * s.head.no)
                s.head.stride = torch.tensor([float(v) for v in strides])
                s.stride, s.names = s.head.stride, ["face"]
                s.maps = None  # {(H, W): [levels (63, ny, nx)]} of the image being processed

            def forward(s, img, augment=False):
                lv = s.maps[(int(img.shape[2]), int(img.shape[3]))]
                return s.head([(l if isinstance(l, torch.Tensor) else torch.from_numpy(l))[None].to(img.device, img.dtype) for l in lv])

        det = msfd.MultiScaleFaceDetector.__new__(msfd.MultiScaleFaceDetector)  # __init__ loads a .pt (msfd:48): none ships
        self.backbone = model if model is not None else _StandInBackbone().eval().to(self.device)
        det.model, det.device = self.backbone, self.device
        det.half = self.device.type != "cpu"                      # msfd:49
        det.stride = int(self.backbone.stride.max())               # msfd:52
        det.img_sizes = [check_img_size(s, s=det.stride) for s in img_sizes]  # msfd:53
        det.names, det.conf_thres, det.iou_thres, det.use_api_preprocess = ["face"], conf_thres, iou_thres, False
        if det.half and model is None:
            self.backbone.half()
        self.det = det

    def detect(self, image, maps, key="mem:0"):
        """detect_multi_scale of one in-memory image. maps: {(H, W): [per-level (63, ny, nx) arrays]}."""
        self.store[key] = image
        if hasattr(self.backbone, "maps"):
            self.backbone.maps = maps
        with _quiet(), self.torch.no_grad():
            final, shape0 = self.det.detect_multi_scale(key)
        return np.asarray(final, np.float64), shape0

    def stage_times(self, image, maps, repeats=3):
        """Per-stage wall clock (best of `repeats`, ms) of the reference's functions on one image: letterbox + pack,
        head decode, non_max_suppression, rescale + round + host copy, weighted_nms."""
        torch, det = self.torch, self.det
        from utils.datasets import letterbox
        from utils.general import non_max_suppression
        from utils.preprocess_yolo_predict import scale_coords_api_approach
        runs = []
        for _ in range(repeats):
            acc = {}

            def clock(name, fn, acc=acc):
                t0 = time.perf_counter()
                r = fn()
                acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
                return r

            all_d = []
            with _quiet(), torch.no_grad():
                for si, size in enumerate(det.img_sizes):
                    def pre():
                        img = letterbox(image, size, stride=det.stride)[0]
                        img = np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))
                        t = torch.from_numpy(img).to(det.device)
                        t = t.half() if det.half else t.float()
                        t /= 255.0
                        return t.unsqueeze(0)
                    x = clock("letterbox_pack", pre)
                    self.backbone.maps = maps
                    pred = clock("head_decode", lambda: self.backbone(x)[0])
                    out = clock("non_max_suppression", lambda: non_max_suppression(pred, det.conf_thres, det.iou_thres))

                    def rescale():
                        rows = []
                        for d in out:
                            if len(d):
                                d[:, :4] = scale_coords_api_approach(x.shape[2:], d[:, :4], image.shape).round()
                                n = d.cpu().numpy()[:, :6]
                                rows.append(np.hstack((n, np.ones((len(n), 1)) * si)))
                        return np.vstack(rows) if rows else np.zeros((0, 7))
                    all_d.append(clock("rescale_round_d2h", rescale))
                merged = np.vstack(all_d)
                clock("weighted_nms", lambda: det.weighted_nms(merged) if len(merged) else merged)
            runs.append(acc)
        return {k: round(1e3 * min(r[k] for r in runs), 3) for k in runs[0]}


def random_init_model(cfg="cfg/yolov7s-face.yaml", seed=0, calib_hw=(256, 384)):
    """A reference network without trained weights that still behaves like one: `models.yolo.Model(cfg)` with random
    init (seeded), then its BatchNorm running statistics are set from three random batches (train-mode forwards, no
    gradient step).  Without that, PyTorch's default conv init lets the activations of the ~100-layer stack die out and
    every head output is its bias (no candidates); with it obj*cls spreads over (0, 1) -- about 1.5 % of the
    rows pass the reference's default conf_thres 0.5 -- and the network is well conditioned, so a last-bit difference in
    its input does not change its output beyond the last bits.  No checkpoint ships with the reference."""
    from baseline import refshim
    if not refshim.install():
        raise RuntimeError("no reference tree")
    import torch
    warnings.filterwarnings("ignore", message="torch.meshgrid")
    with _quiet():
        from models.yolo import Model
        torch.manual_seed(seed)
        net = Model(os.path.join(refshim.REFERENCE_ROOT, cfg))
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = None  # cumulative average over the calibration batches
        m.reset_running_stats()
    net.train()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for _ in range(3):
            net(torch.rand(2, 3, calib_hw[0], calib_hw[1], generator=g))
    for m in bns:
        m.momentum = 0.03  # utils/torch_utils.py:150
    net.eval()
    net.names = ["face"]
    return net


def maps_of(heads, shapes, b):
    """Per-image view of workloads.standin.make_batch output: {(H, W): [levels of image b]}."""
    return {tuple(int(v) for v in shapes[s]): [np.ascontiguousarray(l[b]) for l in heads[s]] for s in range(len(shapes))}


def run_as_written(wl, images, heads, shapes, strides, anchors_px, n_images, harness=None):
    """Mode (i): one process, sequential images and scales, default torch threads.  Returns (seconds, detections)."""
    h = harness or RefHarness(wl["sizes"], wl["conf"], wl["iou"], strides, anchors_px)
    per = [maps_of(heads, shapes, b % len(images)) for b in range(n_images)]
    t0 = time.perf_counter()
    n = 0
    for b in range(n_images):
        final, _ = h.detect(images[b % len(images)], per[b])
        n += len(final)
    return time.perf_counter() - t0, n


def run_port(wl, images, heads, shapes, strides, anchors_px, n_images):
    """Mode (i) with oracle/ref_port.py (no reference tree on this box): the same library calls, one process."""
    from oracle import ref_port
    import torch
    per = [maps_of(heads, shapes, b % len(images)) for b in range(n_images)]
    t0 = time.perf_counter()
    n = 0
    with torch.no_grad():
        for b in range(n_images):
            n += len(ref_port.detect_image(images[b % len(images)], per[b], wl["sizes"], strides, np.asarray(anchors_px, np.float32),
                                           wl["conf"], wl["iou"]))
    return time.perf_counter() - t0, n


# ------------------------------------------------------------------------------------------------ all-cores pool
def _worker_main(conn, wl, strides, anchors_px, per_worker, seed, use_port=False):
    """One single-thread worker process: its own harness and its own shard of synthetic images (spawn start method,
    like yolov7_face_multi_scale_dataframe_predict.py:1115)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    try:
        import torch
        torch.set_num_threads(1)
        from workloads import standin
        images, heads, shapes = standin.make_batch(wl["h"], wl["w"], wl["sizes"], per_worker, wl["faces"], seed=seed,
                                                   radius=wl.get("radius", 1), strides=strides, anchors_px=anchors_px)
        per = [maps_of(heads, shapes, b) for b in range(per_worker)]
        if use_port:
            from oracle import ref_port
            torch.set_grad_enabled(False)

            def detect(b):
                return ref_port.detect_image(images[b], per[b], wl["sizes"], strides, anchors_px, wl["conf"], wl["iou"])
        else:
            h = RefHarness(wl["sizes"], wl["conf"], wl["iou"], strides, anchors_px)

            def detect(b):
                return h.detect(images[b], per[b])[0]
        detect(0)  # page everything in
        conn.send(("ready", 0))
        while True:
            msg = conn.recv()
            if msg == "stop":
                break
            n = 0
            for b in range(per_worker):
                n += len(detect(b))
            conn.send(("done", n))
    except Exception as e:  # surface the failure in the parent instead of hanging it
        conn.send(("error", repr(e)))


class RefPool:
    """`workers` spawned single-thread processes, `per_worker` images each; `step()` = every worker processes its shard
    once (wall clock over all of them)."""

    def __init__(self, wl, strides, anchors_px, workers, per_worker, seed=17, use_port=False):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.workers, self.per_worker = workers, per_worker
        self.procs, self.conns = [], []
        anchors = np.asarray(anchors_px, np.float32)
        for w in range(workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker_main, args=(b, dict(wl), tuple(strides), anchors, per_worker, seed + 101 * w, use_port),
                            daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(a)
        for c in self.conns:
            self._expect(c, "ready")

    @staticmethod
    def _expect(conn, what, timeout=600):
        if not conn.poll(timeout):
            raise RuntimeError("reference worker timed out")
        tag, val = conn.recv()
        if tag != what:
            raise RuntimeError("reference worker failed: %s" % (val,))
        return val

    @property
    def images_per_step(self):
        return self.workers * self.per_worker

    def step(self):
        t0 = time.perf_counter()
        for c in self.conns:
            c.send("go")
        n = sum(self._expect(c, "done") for c in self.conns)
        return time.perf_counter() - t0, n

    def close(self):
        for c in self.conns:
            try:
                c.send("stop")
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()
