#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every output array is produced by the reference's own functions, imported through baseline/refshim.py:
  utils.datasets.letterbox                               (utils/datasets.py:873)
  utils.preprocess_yolo_predict.preprocess_api_approach  (utils/preprocess_yolo_predict.py:345)
  utils.preprocess_yolo_predict.scale_coords_api_approach(utils/preprocess_yolo_predict.py:122)
  models.yolo.Model / IKeypoint.forward                  (models/yolo.py:253-308)
  utils.general.non_max_suppression / scale_coords       (utils/general.py:506, 374)
  multi_scale_face_detector.MultiScaleFaceDetector       (multi_scale_face_detector.py:28-288)
and by torchvision.ops.nms (CPU kernel), the third-party op the reference calls.

The inputs are seeded; large inputs are NOT stored -- the tests regenerate them from the same seeds with
face_detection_multi_scale_b200.synth and compare against stored outputs or SHA-256 digests.
"""
import contextlib
import hashlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

from baseline import refshim  # noqa: E402

assert refshim.install(), "reference tree not found"

import cv2  # noqa: E402
import torch  # noqa: E402
import torchvision  # noqa: E402

from workloads import synth  # noqa: E402
from tests.helpers import (API_CASES, CROWD_HEADS, E2E_CASES, LETTERBOX_FULL, LETTERBOX_SMALL,  # noqa: E402
                           crowd_levels, e2e_heads)

with contextlib.redirect_stdout(io.StringIO()):
    from models.yolo import IKeypoint, Model  # noqa: E402
    from multi_scale_face_detector import MultiScaleFaceDetector  # noqa: E402
    from utils.datasets import letterbox  # noqa: E402
    from utils.general import non_max_suppression, scale_coords  # noqa: E402
    from utils.preprocess_yolo_predict import preprocess_api_approach, scale_coords_api_approach  # noqa: E402

torch.set_grad_enabled(False)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


# ------------------------------------------------------------------------------------------------
# 1. letterbox / pack / api preprocess
# ------------------------------------------------------------------------------------------------


def ref_pack(img_hwc_bgr, half):
    """multi_scale_face_detector.py:96-105 (std branch), on CPU; half=True restates the CUDA dtype path."""
    img = img_hwc_bgr[:, :, ::-1].transpose(2, 0, 1)
    img = np.ascontiguousarray(img)
    t = torch.from_numpy(img)
    t = t.half() if half else t.float()
    t /= 255.0
    return t.numpy()


def gen_letterbox():
    out = {}
    meta = []
    for i, (h, w, ns, auto, scaleup, stride, seed) in enumerate(LETTERBOX_SMALL):
        img = synth.synth_image(h, w, seed)
        o, ratio, pad = letterbox(img, ns, auto=auto, scaleup=scaleup, stride=stride)
        out["small%d_out" % i] = o
        f16, f32 = ref_pack(o, True).view(np.uint16), ref_pack(o, False)
        out["small%d_f16_sha" % i] = np.array(sha(f16))
        out["small%d_f32_sha" % i] = np.array(sha(f32))
        if i in (0, 3):
            out["small%d_f16" % i] = f16
            out["small%d_f32" % i] = f32
        ns2 = ns if isinstance(ns, tuple) else (ns, ns)
        meta.append([h, w, ns2[0], ns2[1], int(auto), int(scaleup), stride, seed, ratio[0], pad[0], pad[1]])
    out["small_meta"] = np.array(meta, np.float64)
    fmeta, fsha_u8, fsha_f16 = [], [], []
    for (h, w, ns, auto, stride, seed) in LETTERBOX_FULL:
        img = synth.synth_image(h, w, seed)
        o, ratio, pad = letterbox(img, ns, auto=auto, stride=stride)
        fmeta.append([h, w, ns, int(auto), stride, seed, o.shape[0], o.shape[1]])
        fsha_u8.append(sha(o))
        fsha_f16.append(sha(ref_pack(o, True)))
    out["full_meta"] = np.array(fmeta, np.int64)
    out["full_sha_u8"] = np.array(fsha_u8)
    out["full_sha_f16"] = np.array(fsha_f16)
    # api approach: needs a file; PNG is lossless.  PIL decodes to RGB.
    ameta = []
    with tempfile.TemporaryDirectory() as td:
        for i, (h, w, size, seed) in enumerate(API_CASES):
            rgb = synth.synth_image(h, w, seed)
            p = os.path.join(td, "a%d.png" % i)
            cv2.imwrite(p, rgb[:, :, ::-1])
            chw = preprocess_api_approach(p, size, 32)
            assert chw.shape == (3, size, size)
            if h * w < 200000:
                out["api%d_out" % i] = chw
            ameta.append([h, w, size, seed])
            out["api%d_sha" % i] = np.array(sha(chw))
    out["api_meta"] = np.array(ameta, np.int64)
    save("letterbox", **out)


# ------------------------------------------------------------------------------------------------
# 2. head decode (+ NMS on the decoded output)
# ------------------------------------------------------------------------------------------------
def identity_head(nl=3, anchors=None):
    """An IKeypoint whose 1x1 convolutions copy their 63-channel input (one-hot weights are exact in fp32),
    so arbitrary synthetic raw head tensors can be pushed through the reference's own decode."""
    if anchors is None:
        anchors = synth.P5_ANCHORS.reshape(nl, -1).tolist()
    with contextlib.redirect_stdout(io.StringIO()):
        m = IKeypoint(nc=1, anchors=anchors, nkpt=5, ch=(63,) * nl)
    for i in range(nl):
        m.ia[i].implicit.zero_()
        m.im[i].implicit.fill_(1.0)
        # IKeypoint concatenates (det 18ch, kpt 45ch) then views as (na, no): channel c of the cat is
        # anchor c // 21, output c % 21 -- NOT "anchor a's det then kpt" (SURVEY App. A4).  Build the
        # permutation so that cat-channel c equals raw-channel c.
        w_all = torch.eye(63).view(63, 63, 1, 1)
        m.m[i].weight.copy_(w_all[:18])
        m.m[i].bias.zero_()
        m.m_kpt[i].weight.copy_(w_all[18:])
        m.m_kpt[i].bias.zero_()
    m.stride = torch.tensor([8., 16., 32., 64.][:nl])
    m.eval()
    return m


def gen_head():
    out = {}
    # (a) real random-init yolov7-tiny-face (cfg/yolov7-tiny-face.yaml), batch 2, 96x160 input
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        net = Model(os.path.join(refshim.REFERENCE_ROOT, "cfg/yolov7-tiny-face.yaml")).eval()
    x = torch.rand(2, 3, 96, 160)
    pred, raw = net(x)
    det = net.model[-1]
    out["tiny_pred"] = pred.numpy()
    for i, r in enumerate(raw):  # (bs, na, ny, nx, no) -> NCHW (bs, na*no, ny, nx)
        bs, na, ny, nx, no = r.shape
        out["tiny_raw%d" % i] = r.permute(0, 1, 4, 2, 3).reshape(bs, na * no, ny, nx).contiguous().numpy()
    out["tiny_strides"] = det.stride.numpy()
    out["tiny_anchors_px"] = det.anchor_grid.view(det.nl, det.na, 2).numpy()
    conf = (pred[..., 4] * pred[..., 5]).flatten()
    thr = float(torch.quantile(conf, 0.90))
    out["tiny_conf_thres"] = np.array(thr)
    for iou in (0.45, 0.3):
        res = quiet(non_max_suppression, pred, thr, iou)
        for b, r in enumerate(res):
            out["tiny_nms_iou%d_b%d" % (int(iou * 100), b)] = r.numpy()

    # (b) synthetic crowd heads pushed through the reference IKeypoint (identity convs)
    head = identity_head()
    for name, (in_h, in_w, n_faces, seed) in CROWD_HEADS.items():
        levels = [torch.from_numpy(l) for l in crowd_levels(in_h, in_w, n_faces, seed)]
        pred, raw = head([l.clone() for l in levels])
        # sanity: identity convs reproduce the input exactly
        for i, r in enumerate(raw):
            bs, na, ny, nx, no = r.shape
            assert torch.equal(r.permute(0, 1, 4, 2, 3).reshape(bs, na * no, ny, nx), levels[i])
        out[name + "_pred_sha"] = np.array(sha(pred.numpy()))
        out[name + "_pred_b0_head"] = pred[0, :64].numpy()
        out[name + "_meta"] = np.array([in_h, in_w, n_faces, seed])
        for conf, iou in ((0.5, 0.5), (0.6, 0.3), (0.25, 0.45)):
            res = quiet(non_max_suppression, pred, conf, iou)
            for b, r in enumerate(res):
                out["%s_nms_c%d_i%d_b%d" % (name, int(conf * 100), int(iou * 100), b)] = r.numpy()
    save("head", **out)


# ------------------------------------------------------------------------------------------------
# 3. torchvision.ops.nms (CPU kernel) known answers
# ------------------------------------------------------------------------------------------------
def gen_nms():
    out = {}
    rng = np.random.default_rng(42)

    def case(name, boxes, scores, thrs):
        boxes = np.ascontiguousarray(boxes, np.float32)
        scores = np.ascontiguousarray(scores, np.float32)
        out[name + "_boxes"] = boxes
        out[name + "_scores"] = scores
        out[name + "_thrs"] = np.array(thrs, np.float64)
        for t in thrs:
            k = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), float(t))
            out["%s_keep_%d" % (name, int(round(t * 1e6)))] = k.numpy()

    # random boxes
    xy = rng.uniform(0, 600, (3000, 2)); wh = rng.uniform(5, 120, (3000, 2))
    case("rand", np.concatenate([xy, xy + wh], 1), rng.uniform(0, 1, 3000), [0.5, 0.3, 0.45])
    # integer boxes (the cross-scale merge sees rounded boxes): many IoUs are small rationals that hit
    # thresholds like 0.3 / 1/3 / 0.5 exactly
    xy = rng.integers(0, 60, (1500, 2)).astype(np.float64); wh = rng.integers(2, 24, (1500, 2))
    case("int", np.concatenate([xy, xy + wh], 1), np.round(rng.uniform(0, 1, 1500), 2), [0.3, 1 / 3, 0.5, 0.6])
    # exact-threshold pairs: IoU == rn32(1/3) and IoU == 0.3f
    case("exact", [[0, 0, 2, 1], [1, 0, 3, 1], [0, 0, 10, 10], [0, 0, 10, 3], [50, 50, 60, 60], [50, 50, 60, 60],
                   [5, 5, 5, 5], [5, 5, 5, 5]],
         [0.9, 0.8, 0.7, 0.6, 0.5, 0.5, 0.4, 0.4], [1 / 3, 0.3, 0.5, 0.0, 1.0])
    # heavy ties + clustered crowd
    c = rng.uniform(50, 550, (40, 2)); sz = rng.uniform(10, 80, 40)
    ci = rng.integers(0, 40, 4000)
    ctr = c[ci] + rng.standard_normal((4000, 2)) * sz[ci, None] * 0.1
    half = sz[ci, None] * np.exp(rng.standard_normal((4000, 2)) * 0.1) / 2
    case("crowd", np.concatenate([ctr - half, ctr + half], 1), np.round(rng.uniform(0.01, 1, 4000), 3), [0.5, 0.3])
    case("one", [[1, 2, 3, 4]], [0.5], [0.5])
    save("nms", **out)


# ------------------------------------------------------------------------------------------------
# 4. rescale
# ------------------------------------------------------------------------------------------------
def gen_rescale():
    out = {}
    rng = np.random.default_rng(5)
    cases = [((384, 640), (1080, 1920, 3)), ((544, 960), (1080, 1920, 3)), ((736, 1280), (1080, 1920, 3)),
             ((640, 640), (480, 270, 3)), ((640, 640), (1080, 1920, 3)), ((896, 1536), (2160, 3840, 3))]
    meta = []
    for i, (in_hw, s0) in enumerate(cases):
        n = 37
        b = rng.uniform(-40, max(in_hw) + 40, (n, 4)).astype(np.float32)
        b[:5] *= 1.5  # push a few out of frame in rows 0..4 (the clip quirk touches rows 0..3 only)
        k = rng.uniform(-40, max(in_hw) + 40, (n, 15)).astype(np.float32)
        out["c%d_boxes" % i] = b
        out["c%d_kpts" % i] = k
        out["c%d_std_boxes" % i] = scale_coords(in_hw, torch.from_numpy(b.copy()), s0).numpy()
        out["c%d_std_kpts" % i] = scale_coords(in_hw, torch.from_numpy(k.copy()), s0, kpt_label=5, step=3).numpy()
        out["c%d_api_boxes" % i] = quiet(scale_coords_api_approach, in_hw, torch.from_numpy(b.copy()), s0).numpy()
        out["c%d_api_round" % i] = quiet(scale_coords_api_approach, in_hw, torch.from_numpy(b.copy()), s0).round().numpy()
        meta.append([in_hw[0], in_hw[1], s0[0], s0[1]])
    out["meta"] = np.array(meta, np.int64)
    save("rescale", **out)


def gen_letterbox_color():
    """letterbox with a per-channel border colour (cv2.copyMakeBorder value=color, utils/datasets.py:902)."""
    out = {}
    cases = [(97, 131, 64, True, (10, 200, 30), 41), (131, 97, 96, False, (0, 255, 114), 42),
             (64, 64, 64, True, (1, 2, 3), 43), (50, 120, (96, 160), False, (114, 114, 7), 44)]
    for i, (h, w, ns, auto, color, seed) in enumerate(cases):
        img = synth.synth_image(h, w, seed)
        o, ratio, pad = letterbox(img, ns, color=color, auto=auto)
        out["c%d_out" % i] = o
        out["c%d_ratio_pad" % i] = np.array([ratio[0], ratio[1], pad[0], pad[1]], np.float64)
    out["meta"] = np.array([[h, w, (ns if isinstance(ns, int) else ns[0]), (ns if isinstance(ns, int) else ns[1]), int(a),
                             c[0], c[1], c[2], seed] for h, w, ns, a, c, seed in cases], np.int64)
    save("letterbox_color", **out)


def gen_rescale_rp():
    """scale_coords with an explicit ratio_pad (utils/general.py:379-383) -- what test.py:172 passes: the dataset's
    shapes = (h0, w0), ((h / h0, w / w0), pad) with pad from letterbox (utils/datasets.py:590-593)."""
    out = {}
    rng = np.random.default_rng(15)
    cases = [((1080, 1920), 640, False), ((480, 270), 640, False), ((2160, 3840), 1536, True), ((333, 517), 960, True),
             ((640, 640), 640, False)]
    meta = []
    for i, ((h0, w0), size, auto) in enumerate(cases):
        img0 = np.zeros((h0, w0, 3), np.uint8)
        img, ratio, pad = letterbox(img0, size, auto=auto, scaleup=True)
        h, w = img.shape[:2]
        for j, rp in enumerate([(ratio, pad), ((h / h0, w / w0), pad), (ratio[0], pad)]):  # tuple gain and scalar gain
            n = 29
            b = rng.uniform(-40, max(h, w) + 40, (n, 4)).astype(np.float32)
            b[:5] *= 1.5
            k = rng.uniform(-40, max(h, w) + 40, (n, 15)).astype(np.float32)
            out["c%d_%d_boxes" % (i, j)] = b
            out["c%d_%d_kpts" % (i, j)] = k
            out["c%d_%d_std_boxes" % (i, j)] = scale_coords((h, w), torch.from_numpy(b.copy()), (h0, w0), rp).numpy()
            out["c%d_%d_std_kpts" % (i, j)] = scale_coords((h, w), torch.from_numpy(k.copy()), (h0, w0), rp,
                                                           kpt_label=5, step=3).numpy()
            g = rp[0][0] if isinstance(rp[0], tuple) else rp[0]
            out["c%d_%d_rp" % (i, j)] = np.array([g, rp[1][0], rp[1][1]], np.float64)
        meta.append([h, w, h0, w0])
    out["meta"] = np.array(meta, np.int64)
    save("rescale_rp", **out)


# ------------------------------------------------------------------------------------------------
# 5. cross-scale merge, 6. end-to-end detect_multi_scale with a stand-in backbone
# ------------------------------------------------------------------------------------------------
def bare_detector(img_sizes, conf, iou, api, model=None):
    d = MultiScaleFaceDetector.__new__(MultiScaleFaceDetector)  # skip attempt_load (no .pt ships)
    d.model = model
    d.device = torch.device("cpu")
    d.half = False
    d.stride = 32
    d.img_sizes = list(img_sizes)
    d.names = ["face"]
    d.conf_thres = conf
    d.iou_thres = iou
    d.use_api_preprocess = api
    return d


def gen_merge():
    out = {}
    rng = np.random.default_rng(9)
    for ci, (n_scales, n, iou) in enumerate([(3, 400, 0.5), (3, 400, 0.3), (2, 250, 0.3), (4, 300, 0.45), (1, 50, 0.5)]):
        det = bare_detector([640] * n_scales, 0.5, iou, False)
        c = rng.uniform(100, 1700, (30, 2)); sz = np.exp(rng.uniform(np.log(12), np.log(300), 30))
        k = rng.integers(0, 30, n)
        ctr = c[k] + rng.standard_normal((n, 2)) * sz[k, None] * 0.06
        half = sz[k, None] * np.exp(rng.standard_normal((n, 2)) * 0.08) / 2
        d = np.zeros((n, 7))
        d[:, :4] = np.round(np.concatenate([ctr - half, ctr + half], 1))
        d[:, 4] = np.round(rng.uniform(0.5, 1, n), 2).astype(np.float32)  # conf came from fp32
        d[:, 6] = np.sort(rng.integers(0, n_scales, n))
        out["m%d_dets" % ci] = d
        out["m%d_weights" % ci] = det.calculate_scale_weights(d)
        out["m%d_out" % ci] = det.weighted_nms(d)
        out["m%d_meta" % ci] = np.array([n_scales, iou])
    save("merge", **out)


class StandInBackbone(torch.nn.Module):
    """Deterministic replacement for the (out-of-scope) backbone: returns, for the input's (H, W), the
    reference IKeypoint decode of pre-generated synthetic raw head tensors.  Everything downstream of the
    head's convolutions is the unmodified reference."""

    def __init__(self, heads_by_hw):
        super().__init__()
        self.head = identity_head()
        self.heads_by_hw = heads_by_hw
        self.seen = {}
        self.stride = torch.tensor([8., 16., 32.])
        self.names = ["face"]

    def forward(self, img, augment=False):
        hw = tuple(img.shape[2:])
        self.seen[hw] = img.clone()
        levels = [torch.from_numpy(l[None].copy()) for l in self.heads_by_hw[hw]]
        return self.head(levels)


def gen_e2e():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for (name, h, w, sizes, n_faces, conf, iou, api, seed) in E2E_CASES:
            bgr = synth.synth_image(h, w, seed)
            p = os.path.join(td, name + ".png")
            cv2.imwrite(p, bgr)
            heads, shapes = e2e_heads(h, w, sizes, n_faces, api, seed)
            model = StandInBackbone(heads)
            det = bare_detector(sizes, conf, iou, api, model)
            final, shape0 = quiet(det.detect_multi_scale, p)
            out[name + "_final"] = np.asarray(final, np.float64)
            out[name + "_shape0"] = np.array(shape0)
            for si, size in enumerate(sizes):
                d, _, _ = quiet(det.detect_single_scale, p, size)
                out["%s_scale%d" % (name, si)] = d
                out["%s_input%d_sha" % (name, si)] = np.array(sha(model.seen[shapes[si]].numpy()))
            out[name + "_shapes"] = np.array(shapes)
    save("e2e", **out)


def gen_nms6():
    """non_max_suppression on a 6-column prediction (kpt_label=0: the no-keypoint single-class head) and on a
    12-column one (kpt_label=2), both cut from the tiny model's decoded output."""
    out = {}
    g = np.load(os.path.join(HERE, "head.npz"))
    pred = torch.from_numpy(g["tiny_pred"])
    thr = float(g["tiny_conf_thres"])
    for cols, kl in ((6, 0), (12, 2)):
        p = pred[..., :cols].contiguous()
        res = quiet(non_max_suppression, p, thr, 0.45, kpt_label=kl)
        for b, r in enumerate(res):
            out["c%d_b%d" % (cols, b)] = r.numpy()
    out["thr"] = np.array(thr)
    save("nms6", **out)


def gen_labels():
    """non_max_suppression with apriori `labels` (autolabelling, utils/general.py:533-541) on the 6-column
    prediction of the tiny model: label boxes [cls, x, y, w, h] partly overlap detections, one image has none."""
    out = {}
    g = np.load(os.path.join(HERE, "head.npz"))
    pred = torch.from_numpy(g["tiny_pred"])[..., :6].contiguous()
    thr = float(g["tiny_conf_thres"])
    base = quiet(non_max_suppression, pred, thr, 0.45, kpt_label=0)
    rng = np.random.default_rng(8)
    labs = []
    for b in range(pred.shape[0]):
        d = base[b].numpy()
        rows = []
        for k in range(min(3, len(d))):  # a label right on top of a detection (suppresses it: conf 1.0 wins)
            x1, y1, x2, y2 = d[k, :4]
            rows.append([0, (x1 + x2) / 2 + rng.uniform(-1, 1), (y1 + y2) / 2, x2 - x1, y2 - y1])
        for k in range(4):  # and labels elsewhere
            rows.append([0, rng.uniform(20, 140), rng.uniform(20, 80), rng.uniform(8, 30), rng.uniform(8, 30)])
        labs.append(torch.tensor(rows, dtype=torch.float32))
    labs[1] = torch.zeros((0, 5))  # second image: no labels
    res = quiet(non_max_suppression, pred, thr, 0.45, kpt_label=0, labels=labs)
    for b, r in enumerate(res):
        out["l%d" % b] = labs[b].numpy()
        out["r%d" % b] = r.numpy()
    out["thr"] = np.array(thr)
    save("labels", **out)


def gen_bbox():
    """widerface_evaluate/box_overlaps.pyx::bbox_overlaps, the reference's own Cython code built by
    `make -C oracle _ref/box_overlaps.so` (np.float was removed from NumPy; the module only needs the alias)."""
    import importlib.machinery
    import importlib.util
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/box_overlaps.so"])
    np.float = float
    loader = importlib.machinery.ExtensionFileLoader("box_overlaps", os.path.join(ROOT, "oracle", "_ref", "box_overlaps.so"))
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("box_overlaps", loader))
    loader.exec_module(mod)
    rng = np.random.default_rng(3)
    out = {}
    for i, (n, k, integer) in enumerate([(300, 40, True), (1000, 7, False), (1, 1, True), (50, 0, True), (0, 5, True)]):
        def boxes(m):
            xy = rng.uniform(0, 400, (m, 2)); wh = rng.uniform(0, 120, (m, 2))
            b = np.concatenate([xy, xy + wh], 1)
            return np.round(b) if integer else b
        a, q = boxes(n), boxes(k)
        out["b%d" % i], out["q%d" % i] = a, q
        out["o%d" % i] = mod.bbox_overlaps(np.ascontiguousarray(a, np.float64), np.ascontiguousarray(q, np.float64))
    save("bbox", **out)


def gen_ncnn():
    """The reference's ncnn-demo NMS (cpp/yolov7-face-ncnn/src/yolov7face.cpp:35-74: std::sort by score, +1 pixel
    extents, `ovr >= thresh`), the reference's own C++ cut out of its source file and built by
    `make -C oracle _ref/ncnn_nms.so`.  Scores are made distinct (std::sort is not stable)."""
    import ctypes as C
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/ncnn_nms.so"])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "ncnn_nms.so"))
    lib.ref_ncnn_nms.restype = C.c_int64
    rng = np.random.default_rng(11)
    out = {}
    cases = [(400, 40, 0.45, False), (1500, 60, 0.3, False), (3000, 25, 0.5, True), (1, 1, 0.5, False), (64, 4, 0.0, True)]
    for i, (n, clusters, thr, integer) in enumerate(cases):
        c = rng.uniform(40, 1200, (clusters, 2))
        sz = rng.uniform(12, 160, (clusters, 2))
        k = rng.integers(0, clusters, n)
        ctr = c[k] + rng.normal(0, 0.08, (n, 2)) * sz[k]
        wh = sz[k] * rng.uniform(0.8, 1.25, (n, 2))
        b = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
        if integer:
            b = np.round(b)  # integer corners: touching boxes exercise the +1 convention and exact ties with thresh
        s = rng.permutation(n).astype(np.float32) / np.float32(n) * np.float32(0.9) + np.float32(0.05)  # distinct
        keep = np.empty((n,), np.int64)
        m = lib.ref_ncnn_nms(b.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_float(thr),
                             keep.ctypes.data_as(C.c_void_p))
        out["b%d" % i], out["s%d" % i], out["t%d" % i], out["k%d" % i] = b, s, np.float32(thr), keep[:m].copy()
    save("ncnn", **out)


def gen_tta():
    """utils/torch_utils.py::scale_img and models/yolo.py::Model._descale_pred / the concatenation of
    forward_augment (:363-374), the reference's own functions on CPU fp32."""
    import types
    from models.yolo import Model
    from utils.torch_utils import scale_img
    torch.manual_seed(3)
    out = {}
    x = torch.rand(2, 3, 48, 80)
    out["x"] = x.numpy()
    for i, (ratio, flip, gs) in enumerate([(0.83, 3, 32), (0.67, None, 32), (0.5, 2, 16), (1.0, 3, 32), (0.83, None, 8)]):
        xi = scale_img(x.flip(flip) if flip else x, ratio, gs=gs)
        out["s%d" % i] = xi.numpy().copy()
        out["s%d_args" % i] = np.array([ratio, flip or 0, gs], np.float64)
    xs = scale_img(x, 0.67, same_shape=True)
    out["same_shape"] = xs.numpy().copy()
    dummy = types.SimpleNamespace(inplace=True)
    p = torch.rand(2, 40, 21) * 300
    out["p"] = p.numpy().copy()
    ys = []
    for i, (si, fi) in enumerate(zip([1, 0.83, 0.67], [None, 3, None])):
        yi = Model._descale_pred(dummy, p.clone(), fi, si, (48, 80))
        out["d%d" % i] = yi.numpy().copy()
        ys.append(yi)
    out["d_ud"] = Model._descale_pred(dummy, p.clone(), 2, 0.83, (48, 80)).numpy().copy()
    out["cat"] = torch.cat(ys, 1).numpy().copy()
    save("tta", **out)


def gen_pack():
    """normalize_bbox + create_yolo_json_format (utils/preprocess_yolo_predict.py:26-49,478-571) on synthetic merged
    detections, assembled per frame exactly like yolov7_face_multi_scale_dataframe_predict.py:699-716."""
    from utils.preprocess_yolo_predict import create_yolo_json_format, normalize_bbox
    rng = np.random.default_rng(21)
    W, H = 1920, 1080
    counts = [5, 0, 12, 1]
    frames, dets_all = [], []
    for fi, n in enumerate(counts):
        xy = np.round(rng.uniform(0, 1000, (n, 2))); wh = np.round(rng.uniform(5, 300, (n, 2)))
        d = np.zeros((n, 7), np.float32)
        d[:, :2], d[:, 2:4] = xy, xy + wh
        d[:, 4] = rng.uniform(0.5, 1, n).astype(np.float32)
        d[:, 6] = rng.integers(0, 3, n)
        dets_all.append(d)
        fd = {"num_faces": n, "bboxes": [], "confidence": [], "class_names": [], "class_indexes": [], "class_groups": [],
              "infer_time": 0.01 * (fi + 1)}
        for r in d:  # :708-716
            fd["bboxes"].append(normalize_bbox([float(v) for v in r[:4]], W, H))
            fd["confidence"].append(float(r[4]))
            fd["class_names"].append("face"); fd["class_indexes"].append(0); fd["class_groups"].append("face")
        frames.append(fd)
    js = create_yolo_json_format(frames)["yolo_face_prediction"]
    out = {"counts": np.array(counts), "wh": np.array([W, H])}
    for fi, d in enumerate(dets_all):
        out["dets%d" % fi] = d
    out["bboxes"] = np.array(js[0]["data"], np.float64)
    out["confidence"] = np.array(js[1]["data"], np.float64)
    out["class_indexes"] = np.array(js[3]["data"], np.int32)
    out["shape"] = np.array(js[0]["shape"])
    save("pack", **out)


def gen_ikeypoint():
    """A REAL reference IKeypoint head (models/yolo.py:215-313: ImplicitA -> 1x1 conv -> ImplicitM, plus the
    keypoint conv, cat, view/permute, decode) with seeded random weights on small feature maps.  The weights, the
    inputs and the reference's outputs are stored so that the patched head (yolo_head.patch_head on a structurally
    identical module) can be checked against the reference's own forward on the GPU box."""
    torch.manual_seed(7)
    ch = (8, 16, 32)
    with contextlib.redirect_stdout(io.StringIO()):
        m = IKeypoint(nc=1, anchors=synth.P5_ANCHORS.reshape(3, -1).tolist(), nkpt=5, ch=ch)
    m.stride = torch.tensor([8., 16., 32.])
    m.eval()
    for p_ in m.parameters():  # spread the logits so that the confidence filter sees both sides
        p_.data.mul_(3.0)
    out = {}
    for k, v in m.state_dict().items():
        out["w_" + k.replace(".", "__")] = v.numpy()
    xs = [torch.randn(2, c, 96 // s, 160 // s) for c, s in zip(ch, (8, 16, 32))]
    for i, x in enumerate(xs):
        out["x%d" % i] = x.numpy()
    pred, raw = m([x.clone() for x in xs])
    out["pred"] = pred.numpy()
    for i, r in enumerate(raw):
        out["raw%d" % i] = r.numpy()
    conf = (pred[..., 4] * pred[..., 5]).flatten()
    thr = float(torch.quantile(conf, 0.97))
    out["thr"] = np.array(thr)
    for b, r in enumerate(quiet(non_max_suppression, pred, thr, 0.45)):
        out["nms_b%d" % b] = r.numpy()
    save("ikeypoint", **out)


# ------------------------------------------------------------------------------------------------
# 16. every head variant the reference can run in eval mode (Detect / IDetect / IKeypoint, any na / nc / nkpt)
# ------------------------------------------------------------------------------------------------
HEAD_VARIANTS = {
    # name: (class, nc, nkpt, na, inplace)
    "detect_nc3": ("Detect", 3, 0, 3, False),       # plain multi-class head, non-inplace branch (models/yolo.py:95-99)
    "detect_nc80": ("Detect", 80, 0, 3, False),     # COCO-sized rows (no = 85)
    "idetect_nc1": ("IDetect", 1, 0, 2, True),      # no = 6 through the inplace branch, two anchors per level
    "idetect_pose": ("IDetect", 1, 17, 3, True),    # 17 keypoints (no = 57): the keypoint columns of x are decoded in place
    "detect_face": ("Detect", 1, 5, 3, True),       # the face layout on a Detect head (x mutated, :76-81)
    "ikeypoint_k3": ("IKeypoint", 1, 3, 4, True),   # three keypoints, four anchors per level (no = 15)
}


def gen_heads():
    import models.yolo as Y
    out = {}
    ch, strides, hw = (8, 16, 32), (8., 16., 32.), (32, 64)
    for vi, (name, (cls, nc, nkpt, na, inplace)) in enumerate(sorted(HEAD_VARIANTS.items())):
        torch.manual_seed(100 + vi)
        rng = np.random.default_rng(100 + vi)
        anchors = (rng.uniform(4, 200, size=(3, na * 2))).round(1).tolist()
        with contextlib.redirect_stdout(io.StringIO()):
            m = getattr(Y, cls)(nc=nc, anchors=anchors, nkpt=nkpt, ch=ch, inplace=inplace)
        m.stride = torch.tensor(strides)
        m.eval()
        for p_ in m.parameters():
            p_.data.mul_(3.0)
        with_module = name != "detect_nc80"  # weights + conv inputs for the patched-module test (kept small)
        if with_module:
            for k, v in m.state_dict().items():
                out["%s__w_%s" % (name, k.replace(".", "__"))] = v.numpy()
        xs = [torch.randn(2, c, int(hw[0] // s), int(hw[1] // s)) for c, s in zip(ch, strides)]
        # the decode's input: the conv outputs exactly as the head's forward computes them (:62-65, :155-158, :270-273)
        for i, x in enumerate(xs):
            det = m.im[i](m.m[i](m.ia[i](x))) if hasattr(m, "ia") else m.m[i](x)
            raw = det if not nkpt else torch.cat((det, m.m_kpt[i](x)), 1)
            if with_module:
                out["%s__x%d" % (name, i)] = x.numpy()
            out["%s__raw%d" % (name, i)] = raw.numpy()
        pred, xo = m([x.clone() for x in xs])
        assert pred.shape[-1] == m.no
        out[name + "__pred"] = pred.numpy()
        for i, r in enumerate(xo):  # second return value: (bs, na, ny, nx, no); Detect / IDetect decode its keypoint columns
            bs, na_, ny, nx, no = r.shape
            raw_view = torch.from_numpy(out["%s__raw%d" % (name, i)]).view(bs, na_, no, ny, nx).permute(0, 1, 3, 4, 2)
            if cls == "IKeypoint" or not nkpt:
                assert torch.equal(r, raw_view)  # un-mutated
            else:
                assert torch.equal(r[..., :6], raw_view[..., :6]) and not torch.equal(r[..., 6:], raw_view[..., 6:])
                out["%s__xo%d" % (name, i)] = r.numpy()
        out[name + "__meta"] = np.array([nc, nkpt, na, int(inplace)])
        out[name + "__anchors_px"] = m.anchor_grid.view(m.nl, m.na, 2).numpy()
    save("heads", **out)


def gen_heads_half():
    """The same head classes as HALF models (the reference's default on a GPU, msfd:49), run here on CPU half tensors: ATen
    rounds every half operation the same way on either device (opmath float32, result stored as half); only the exponential
    inside the sigmoid differs in the last float32 bit, which the GPU check (tests/ref_gpu_check.py P1h) pins.  Stored: the
    half conv outputs and the reference's float32 pred."""
    import models.yolo as Y
    out = {}
    ch, strides, hw = (8, 16, 32), (8., 16., 32.), (32, 64)
    for vi, name in enumerate(("detect_face", "idetect_pose", "ikeypoint_k3", "idetect_nc1")):
        cls, nc, nkpt, na, inplace = HEAD_VARIANTS[name]
        torch.manual_seed(200 + vi)
        rng = np.random.default_rng(200 + vi)
        anchors = (rng.uniform(4, 300, size=(3, na * 2))).round(2).tolist()
        with contextlib.redirect_stdout(io.StringIO()):
            m = getattr(Y, cls)(nc=nc, anchors=anchors, nkpt=nkpt, ch=ch, inplace=inplace)
        m.stride = torch.tensor(strides)
        m.eval()
        for p_ in m.parameters():
            p_.data.mul_(3.0)
        m.half()
        xs = [torch.randn(2, c, int(hw[0] // s), int(hw[1] // s)).half() for c, s in zip(ch, strides)]
        for i, x in enumerate(xs):
            det = m.im[i](m.m[i](m.ia[i](x))) if hasattr(m, "ia") else m.m[i](x)
            raw = det if not nkpt else torch.cat((det, m.m_kpt[i](x)), 1)
            assert raw.dtype == torch.float16
            out["%s__raw%d" % (name, i)] = raw.numpy()
        pred, _ = m([x.clone() for x in xs])
        assert pred.dtype == torch.float32  # the grid is float32: `+ grid` and torch.cat promote (models/yolo.py:290-297)
        out[name + "__pred"] = pred.numpy()
        out[name + "__meta"] = np.array([nc, nkpt, na, int(inplace)])
        out[name + "__anchors_px"] = m.anchor_grid.view(m.nl, m.na, 2).float().numpy()
    save("heads_half", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        globals()["gen_" + sys.argv[1]]()
        sys.exit(0)
    gen_letterbox()
    gen_head()
    gen_nms()
    gen_rescale()
    gen_rescale_rp()
    gen_letterbox_color()
    gen_merge()
    gen_e2e()
    gen_nms6()
    gen_labels()
    gen_bbox()
    gen_ncnn()
    gen_tta()
    gen_pack()
    gen_ikeypoint()
    gen_heads()
    gen_heads_half()
