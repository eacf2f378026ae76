#!/usr/bin/env python
"""Randomised differential check of the CPU oracle against the UNMODIFIED reference, run live.

Only where the reference tree exists (the build container): the reference is imported through baseline/refshim.py in
THIS process (tests/test_oracle_golden.py starts it as a subprocess, so the reference's top-level `utils` / `models`
packages never enter the test process).  The committed fixtures pin the oracle on fixed cases; this pins it on
seeded random ones, function by function:

    utils.datasets.letterbox                                  (utils/datasets.py:873)    bit exact
    utils.preprocess_yolo_predict.preprocess_api_approach     (ppyp:345)                 bit exact
    models.yolo.IKeypoint.forward, inference branch           (models/yolo.py:267-308)   <= 1e-4 px or 4 ulp (ATen's expf)
    utils.general.non_max_suppression                         (utils/general.py:506)     bit exact on equal inputs
    utils.general.scale_coords (with and without ratio_pad)   (utils/general.py:374)     bit exact
    utils.preprocess_yolo_predict.scale_coords_api_approach   (ppyp:122)                 bit exact
    MultiScaleFaceDetector.calculate_scale_weights / weighted_nms (msfd:168-240)         bit exact
    MultiScaleFaceDetector.detect_multi_scale around a stand-in backbone (msfd:242-288)  bit exact
    utils.torch_utils.scale_img, Model._descale_pred          (torch_utils.py:247, yolo.py:402)  <= 2 ulp / bit exact
    models.yolo.Detect / IDetect / IKeypoint, any na / nc / nkpt (models/yolo.py:59-308)  <= 1e-4 px or 4 ulp; refusals
    the same heads as HALF models (msfd:49): float32 pred with half-rounded columns        bit exact

Prints one JSON line {"checks": n, "failures": [...]} and exits 1 if anything differs.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

from oracle import oracle as O  # noqa: E402
from baseline import refshim  # noqa: E402

if not refshim.install():
    print(json.dumps({"checks": 0, "failures": [], "skipped": "no reference tree"}))
    sys.exit(0)

import cv2  # noqa: E402
import torch  # noqa: E402

from workloads import synth  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from models.yolo import IKeypoint, Model  # noqa: E402
    from multi_scale_face_detector import MultiScaleFaceDetector  # noqa: E402
    from utils.datasets import letterbox  # noqa: E402
    from utils.general import non_max_suppression, scale_coords  # noqa: E402
    from utils.preprocess_yolo_predict import preprocess_api_approach, scale_coords_api_approach  # noqa: E402
    from utils.torch_utils import scale_img  # noqa: E402

torch.set_grad_enabled(False)
FAIL, N = [], [0]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def check(ok, what):
    N[0] += 1
    if not ok:
        FAIL.append(what)


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    rng = np.random.default_rng(9000 + seed)

    # ---- letterbox: sizes, auto / scaleFill / scaleup, strides, colours
    for it in range(30):
        h, w = int(rng.integers(5, 400)), int(rng.integers(5, 400))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ns = int(rng.integers(16, 420)) if it % 3 else (int(rng.integers(16, 300)), int(rng.integers(16, 300)))
        auto, fill, up = bool(rng.integers(0, 2)), bool(it % 7 == 0), bool(rng.integers(0, 2))
        stride = int(rng.choice([32, 64]))
        color = (114, 114, 114) if it % 4 else tuple(int(c) for c in rng.integers(0, 256, 3))
        try:
            ref, rr, rp = letterbox(img, ns, color=color, auto=auto, scaleFill=fill, scaleup=up, stride=stride)
        except cv2.error:
            continue  # degenerate target (a side rounds to 0): the reference itself fails
        got, gr, gp = O.letterbox(img, ns, color=color, auto=auto, scaleFill=fill, scaleup=up, stride=stride)
        check(got.shape == ref.shape and np.array_equal(got, ref) and tuple(gr) == tuple(rr) and tuple(gp) == tuple(rp),
              ("letterbox", it, h, w, ns, auto, fill, up, stride, color))

    # ---- api preprocessing (PIL decode of a lossless file -> zero-padded square -> letterbox(auto=False) -> CHW)
    with tempfile.TemporaryDirectory() as d:
        for it in range(8):
            h, w = int(rng.integers(20, 300)), int(rng.integers(20, 300))
            rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            path = os.path.join(d, "a%d.png" % it)
            cv2.imwrite(path, rgb[:, :, ::-1])
            size = int(rng.integers(1, 10)) * 32
            ref = quiet(preprocess_api_approach, path, size, 32)
            ref = ref[0] if isinstance(ref, tuple) else ref
            check(np.array_equal(O.preprocess_api(rgb, size), np.asarray(ref)), ("preprocess_api", it, h, w, size))

    # ---- head decode: a real IKeypoint with random weights on random maps
    torch.manual_seed(seed)
    ch = (8, 16, 32)
    with contextlib.redirect_stdout(io.StringIO()):
        m = IKeypoint(nc=1, anchors=synth.P5_ANCHORS.reshape(3, -1).tolist(), nkpt=5, ch=ch)
    m.stride = torch.tensor([8., 16., 32.])
    m.eval()
    for p_ in m.parameters():
        p_.data.mul_(3.0)
    for it in range(4):
        ih, iw = int(rng.integers(1, 6)) * 32, int(rng.integers(1, 6)) * 32
        xs = [torch.randn(2, c, ih // s, iw // s) for c, s in zip(ch, (8, 16, 32))]
        pred, raw = m([x.clone() for x in xs])
        levels = []
        for r in raw:  # (bs, na, ny, nx, no) -> NCHW
            bs, na, ny, nx, no = r.shape
            levels.append(r.permute(0, 1, 4, 2, 3).reshape(bs, na * no, ny, nx).contiguous().numpy())
        anchors_px = m.anchor_grid.view(m.nl, m.na, 2).numpy()
        mine = O.decode_head(levels, (8, 16, 32), anchors_px)
        # ATen's vectorised expf (Sleef, 1 ulp) against the correctly rounded one: <= 1e-4 px, or -- for box sizes
        # above 512 px, where one fp32 ulp is already 6.1e-5 px -- within 4 ulp of the value
        ref_np = pred.numpy()
        d = np.abs(mine - ref_np) if mine.shape == ref_np.shape else np.array([np.inf])
        ulps = np.abs(mine.view(np.int32).astype(np.int64) - ref_np.view(np.int32).astype(np.int64)) \
            if mine.shape == ref_np.shape else np.array([1 << 40])
        check(bool(np.all((d <= 1e-4) | (ulps <= 4))) and float(np.mean(ulps == 0)) > 0.98, ("decode", it, ih, iw, float(d.max())))
        # NMS on the reference's own pred: equal inputs -> identical outputs
        conf = float(np.quantile((pred[..., 4] * pred[..., 5]).numpy(), 0.9))
        for iou in (0.3, 0.45):
            ref = quiet(non_max_suppression, pred, conf, iou)
            got = O.non_max_suppression(pred.numpy(), conf, iou, O.TV_CPU)
            check(all(g.shape == tuple(r.shape) and np.array_equal(g, r.numpy()) for g, r in zip(got, ref)),
                  ("nms_on_ref_pred", it, iou))

    # ---- non_max_suppression on clustered synthetic predictions (ties, many suppressions)
    for it in range(10):
        rows = int(rng.choice([1, 50, 800, 4000]))
        pred = np.stack([synth.synth_pred(rows, int(rng.integers(1, 15)), int(rng.integers(1, 50)), 50 * it + b,
                                          tie_decimals=2 if it % 3 == 0 else None) for b in range(2)])
        conf, iou = float(rng.choice([0.1, 0.3, 0.5])), float(rng.choice([0.3, 0.45, 0.6]))
        ref = quiet(non_max_suppression, torch.from_numpy(pred), conf, iou)
        got = O.non_max_suppression(pred, conf, iou, O.TV_CPU)
        check(all(g.shape == tuple(r.shape) and np.array_equal(g, r.numpy()) for g, r in zip(got, ref)),
              ("nms_synth", it, rows, conf, iou))

    # ---- rescale
    for it in range(20):
        ih, iw = int(rng.integers(2, 40)) * 32, int(rng.integers(2, 40)) * 32
        h0, w0 = int(rng.integers(30, 3000)), int(rng.integers(30, 3000))
        n = int(rng.choice([1, 3, 4, 5, 60]))
        b = rng.uniform(-50, max(ih, iw) + 50, (n, 4)).astype(np.float32)
        k = rng.uniform(-50, max(ih, iw) + 50, (n, 15)).astype(np.float32)
        rp = None if it % 2 else ((float(rng.uniform(0.1, 2.0)),) * 2, (float(rng.uniform(0, 40)), float(rng.uniform(0, 40))))
        check(np.array_equal(O.scale_coords((ih, iw), b.copy(), (h0, w0), rp),
                             scale_coords((ih, iw), torch.from_numpy(b.copy()), (h0, w0), rp).numpy()), ("scale_coords", it))
        check(np.array_equal(O.scale_coords((ih, iw), k.copy(), (h0, w0), rp, kpt_label=5, step=3),
                             scale_coords((ih, iw), torch.from_numpy(k.copy()), (h0, w0), rp, kpt_label=5, step=3).numpy()),
              ("scale_coords_kpt", it))
        ref = quiet(scale_coords_api_approach, (ih, iw), torch.from_numpy(b.copy()), (h0, w0, 3))
        check(np.array_equal(O.scale_coords_api((ih, iw), b.copy(), (h0, w0)), ref.numpy()), ("scale_coords_api", it))
        check(np.array_equal(O.scale_coords_api((ih, iw), b.copy(), (h0, w0), do_round=True), ref.round().numpy()),
              ("scale_coords_api_round", it))

    # ---- cross-scale merge
    for it in range(8):
        n_scales = int(rng.integers(1, 5))
        det = MultiScaleFaceDetector.__new__(MultiScaleFaceDetector)
        det.img_sizes, det.iou_thres = [640] * n_scales, float(rng.choice([0.3, 0.45, 0.5]))
        n = int(rng.integers(1, 500))
        c = rng.uniform(100, 1700, (30, 2)); sz = np.exp(rng.uniform(np.log(12), np.log(300), 30))
        kk = rng.integers(0, 30, n)
        ctr = c[kk] + rng.standard_normal((n, 2)) * sz[kk, None] * 0.06
        half = sz[kk, None] * np.exp(rng.standard_normal((n, 2)) * 0.08) / 2
        d = np.zeros((n, 7))
        d[:, :4] = np.round(np.concatenate([ctr - half, ctr + half], 1))
        d[:, 4] = np.round(rng.uniform(0.3, 1.0, n), 3 if it % 2 else 7)
        d[:, 6] = rng.integers(0, n_scales, n)
        check(np.array_equal(O.scale_weights(d, n_scales), quiet(det.calculate_scale_weights, d)), ("scale_weights", it))
        check(np.array_equal(O.weighted_nms(d, n_scales, det.iou_thres), quiet(det.weighted_nms, d)), ("weighted_nms", it))

    # ---- the whole detect_multi_scale (msfd:242-288) around a stand-in backbone: an IKeypoint whose 1x1 convs copy
    # their input, fed with synthetic raw head maps chosen by the input's (H, W); everything after the convolutions
    # is the unmodified reference
    from tests.helpers import e2e_heads

    def identity_head():
        with contextlib.redirect_stdout(io.StringIO()):
            hd = IKeypoint(nc=1, anchors=synth.P5_ANCHORS.reshape(3, -1).tolist(), nkpt=5, ch=(63,) * 3)
        eye = torch.eye(63).view(63, 63, 1, 1)
        for i in range(3):
            hd.ia[i].implicit.zero_()
            hd.im[i].implicit.fill_(1.0)
            hd.m[i].weight.copy_(eye[:18]); hd.m[i].bias.zero_()
            hd.m_kpt[i].weight.copy_(eye[18:]); hd.m_kpt[i].bias.zero_()
        hd.stride = torch.tensor([8., 16., 32.])
        return hd.eval()

    class StandIn(torch.nn.Module):
        def __init__(self, heads_by_hw):
            super().__init__()
            self.head, self.heads_by_hw = identity_head(), heads_by_hw
            self.stride, self.names = torch.tensor([8., 16., 32.]), ["face"]

        def forward(self, img, augment=False):
            return self.head([torch.from_numpy(l[None].copy()) for l in self.heads_by_hw[tuple(img.shape[2:])]])

    with tempfile.TemporaryDirectory() as d:
        for it in range(5):
            h, w = int(rng.integers(120, 500)), int(rng.integers(120, 500))
            sizes = sorted(set(int(v) * 32 for v in rng.integers(3, 12, int(rng.integers(1, 4)))))
            api = bool(it & 1)
            conf, iou = float(rng.choice([0.4, 0.5, 0.6])), float(rng.choice([0.3, 0.5]))
            n_faces = int(rng.integers(0, 30))
            heads, shapes = e2e_heads(h, w, sizes, n_faces, api, 700 + 10 * seed + it)
            path = os.path.join(d, "e%d.png" % it)
            cv2.imwrite(path, synth.synth_image(h, w, it))
            det = MultiScaleFaceDetector.__new__(MultiScaleFaceDetector)
            det.model, det.device, det.half, det.stride = StandIn(heads), torch.device("cpu"), False, 32
            det.img_sizes, det.names, det.conf_thres, det.iou_thres = list(sizes), ["face"], conf, iou
            det.use_api_preprocess = api
            ref, shape0 = quiet(det.detect_multi_scale, path)
            ref = np.asarray(ref, np.float64)
            got, _ = O.detect_from_heads([[l[None] for l in heads[s_]] for s_ in shapes], shapes, (h, w, 3),
                                         synth.P5_STRIDES, synth.P5_ANCHORS, conf, iou, O.TV_CPU)
            check(tuple(shape0) == (h, w, 3) and got.shape == ref.shape and (ref.size == 0 or np.array_equal(got, ref)),
                  ("detect_multi_scale", it, h, w, sizes, api, conf, iou, n_faces, got.shape, ref.shape))
            # the same call as the reference makes it on a GPU: a half model fed half images (msfd:48-49,101); float32 pred
            heads16 = {k: [l.astype(np.float16) for l in v] for k, v in heads.items()}
            det.model, det.half = StandIn(heads16).half(), True
            ref, _ = quiet(det.detect_multi_scale, path)
            ref = np.asarray(ref, np.float64)
            got, _ = O.detect_from_heads([[l[None] for l in heads16[s_]] for s_ in shapes], shapes, (h, w, 3),
                                         synth.P5_STRIDES, synth.P5_ANCHORS, conf, iou, O.TV_CPU, half_model=True)
            check(got.shape == ref.shape and (ref.size == 0 or np.array_equal(got, ref)),
                  ("detect_multi_scale, half model", it, h, w, sizes, api, conf, iou, n_faces, got.shape, ref.shape))

    # ---- source discovery of the inference loader (utils/datasets.py:126-141): directory, glob, single file
    from utils.datasets import LoadImages as RefLoadImages
    from face_detection_multi_scale_b200 import datasets as D
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "sub"))
        for name in ("b.jpg", "a.PNG", "notes.txt", "z.mp4", "c.jpeg", "sub/e.bmp", "m.avi", "k.webp"):
            open(os.path.join(d, name), "wb").close()
        for spec in (d, os.path.join(d, "*.jp*"), os.path.join(d, "**", "*.*"), os.path.join(d, "b.jpg")):
            ref = quiet(RefLoadImages, spec, 64, 32)
            found = D._expand_source(spec)
            stills = [f for f in found if D._suffix(f) in D.IMG_FORMATS]
            clips = [f for f in found if D._suffix(f) in D.VID_FORMATS]
            check(stills + clips == list(ref.files) and [False] * len(stills) + [True] * len(clips) == list(ref.video_flag)
                  and len(stills + clips) == ref.nf, ("LoadImages discovery", spec.replace(d, "<tmp>")))

    # ---- stream loader: source list parsing, names, rect decision, first tick (utils/datasets.py:262-340), without a GPU:
    # the drop-in's host logic against the reference's on two file "cameras"
    from utils.datasets import LoadStreams as RefLoadStreams
    with tempfile.TemporaryDirectory() as d:
        for name, (hh, ww) in (("a.avi", (90, 120)), ("b.avi", (60, 120))):
            vw = cv2.VideoWriter(os.path.join(d, name), cv2.VideoWriter_fourcc(*"MJPG"), 25, (ww, hh))
            for i in range(20):
                vw.write(rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8))
            vw.release()
        lst = os.path.join(d, "streams.txt")
        for names in (["a.avi"], ["a.avi", "b.avi"], ["a.avi", "a.avi"]):
            with open(lst, "w") as f:
                f.write("\n".join(os.path.join(d, n) for n in names) + "\n")
            try:
                ref = quiet(RefLoadStreams, lst, 64, 32)
            except Exception:
                continue
            geoms = [O.letterbox_geometry(x.shape[0], x.shape[1], 64, auto=True, stride=32) for x in ref.imgs]
            rect = len({(g.dst_h, g.dst_w) for g in geoms}) == 1
            check(rect == bool(ref.rect) and [D.clean_str(os.path.join(d, n)) for n in names] == list(ref.sources),
                  ("LoadStreams rect / names", names))

    # ---- test-time augmentation pieces
    for it in range(6):
        shape = (int(rng.integers(1, 3)), 3, int(rng.integers(1, 8)) * 32, int(rng.integers(1, 8)) * 32)
        x = rng.random(shape, dtype=np.float32)
        ratio = float(rng.choice([0.83, 0.67, 0.5]))
        ref = scale_img(torch.from_numpy(x), ratio, gs=32).numpy()
        got = O.scale_img(x, ratio, gs=32)
        check(got.shape == ref.shape and float(np.max(np.abs(got - ref))) <= 2.4e-7, ("scale_img", it, shape, ratio))
        p = rng.uniform(0, 300, (shape[0], 40, 21)).astype(np.float32)
        flips = [None, 2, 3][it % 3]
        ref = Model._descale_pred(type("M", (), {"inplace": True})(), torch.from_numpy(p.copy()), flips, ratio,
                                  shape[2:]).numpy()
        check(np.array_equal(O.descale_pred(p.copy(), flips, ratio, shape[2:]), ref), ("descale_pred", it))

    # ---- every head class of models/yolo.py, random geometry: the oracle's general decode against the live modules, and the
    # combinations face_detection_multi_scale_b200.yolo_head.head_columns refuses really fail (or scramble) in the reference
    import models.yolo as Y
    for it in range(12):
        cls = ("Detect", "IDetect", "IKeypoint")[it % 3]
        nkpt = int(rng.choice([0, 1, 5, 17]))
        nc, inplace = (1, True) if nkpt else (int(rng.choice([1, 2, 20])), False)
        if nkpt == 0 and it % 2:
            nc, inplace = 1, True  # no = 6 through the inplace branch
        na, nl = int(rng.integers(1, 5)), int(rng.integers(1, 4))
        anchors = rng.uniform(4, 150, size=(nl, na * 2)).round(1).tolist()
        ch = tuple(int(c) for c in rng.integers(2, 9, nl))
        m = quiet(getattr(Y, cls), nc=nc, anchors=anchors, nkpt=nkpt, ch=ch, inplace=inplace)
        strides = [8.0 * 2 ** i for i in range(nl)]
        m.stride = torch.tensor(strides)
        m.eval()
        for p_ in m.parameters():
            p_.data.mul_(3.0)
        hw = (int(rng.integers(1, 4)) * 8 * 2 ** (nl - 1), int(rng.integers(1, 4)) * 8 * 2 ** (nl - 1))
        xs = [torch.from_numpy(rng.normal(size=(2, c, int(hw[0] // s), int(hw[1] // s))).astype(np.float32))
              for c, s in zip(ch, strides)]
        raws = []
        for i, x in enumerate(xs):
            det = m.im[i](m.m[i](m.ia[i](x))) if hasattr(m, "ia") else m.m[i](x)
            raws.append((det if not nkpt else torch.cat((det, m.m_kpt[i](x)), 1)).numpy())
        ref = m([x.clone() for x in xs])[0].numpy()
        got = O.decode_head_general(raws, strides, m.anchor_grid.view(nl, na, 2).numpy(), na, m.no, nkpt)
        check(got.shape == ref.shape and bool(np.all(np.abs(got - ref) <= np.maximum(1e-4, 4 * np.spacing(np.abs(ref))))),
              ("head variant", cls, nc, nkpt, na, nl, inplace))
    for cls, nc, nkpt, inplace in (("Detect", 2, 5, True), ("IDetect", 1, 5, False), ("IKeypoint", 1, 5, False),
                                   ("IKeypoint", 2, 5, True), ("Detect", 80, 0, True), ("IDetect", 3, 0, True)):
        m = quiet(getattr(Y, cls), nc=nc, anchors=[[10, 13, 16, 30, 33, 23]], nkpt=nkpt, ch=(8,), inplace=inplace)
        m.stride = torch.tensor([8.])
        m.eval()
        try:
            rows = m([torch.zeros(1, 8, 4, 6)])[0].shape[1]
            check(rows != 3 * 4 * 6, ("the reference runs a head that head_columns refuses", cls, nc, nkpt, inplace))
        except RuntimeError:
            check(True, "refused combination raises in the reference")

    # ---- a HALF model (the reference's default on a GPU): pred is float32 with half-rounded columns; the oracle's
    # restatement of where ATen rounds against the live modules on half tensors
    for it in range(6):
        cls = ("Detect", "IDetect", "IKeypoint")[it % 3]
        nkpt = [5, 5, 5, 0, 17, 1][it]
        na, nl = int(rng.integers(1, 5)), int(rng.integers(1, 4))
        anchors = rng.uniform(4, 300, size=(nl, na * 2)).round(2).tolist()
        ch = tuple(int(c) for c in rng.integers(2, 9, nl))
        m = quiet(getattr(Y, cls), nc=1, anchors=anchors, nkpt=nkpt, ch=ch, inplace=True)
        strides = [8.0 * 2 ** i for i in range(nl)]
        m.stride = torch.tensor(strides)
        m.eval()
        for p_ in m.parameters():
            p_.data.mul_(3.0)
        m.half()
        hw = (int(rng.integers(1, 4)) * 8 * 2 ** (nl - 1), int(rng.integers(1, 4)) * 8 * 2 ** (nl - 1))
        xs = [torch.from_numpy(rng.normal(size=(2, c, int(hw[0] // s), int(hw[1] // s))).astype(np.float16))
              for c, s in zip(ch, strides)]
        raws = []
        for i, x in enumerate(xs):
            det = m.im[i](m.m[i](m.ia[i](x))) if hasattr(m, "ia") else m.m[i](x)
            raws.append((det if not nkpt else torch.cat((det, m.m_kpt[i](x)), 1)).numpy())
        ref = m([x.clone() for x in xs])[0]
        got = O.decode_head_half(raws, strides, m.anchor_grid.view(nl, na, 2).float().numpy(), na, m.no, nkpt)
        check(ref.dtype == torch.float32 and got.shape == tuple(ref.shape) and np.array_equal(got, ref.numpy()),
              ("half-model head", cls, nkpt, na, nl, float(np.max(np.abs(got - ref.numpy()))) if got.shape == tuple(ref.shape) else None))

    print(json.dumps({"checks": N[0], "failures": [repr(f) for f in FAIL[:20]]}))
    sys.stdout.flush()
    # leave without interpreter finalisation: the reference's LoadStreams reader threads and the CPU half kernels' worker
    # threads are still alive and abort in their destructors ("terminate called without an active exception")
    os._exit(1 if FAIL else 0)


if __name__ == "__main__":
    main()
