"""The reference-facing Python surface (drop-in signatures of utils/datasets.py, utils/general.py,
utils/preprocess_yolo_predict.py, models/yolo.py head, multi_scale_face_detector.py) on the CUDA path, against
golden outputs of the UNMODIFIED reference and against the oracle.  GPU only."""
import hashlib
import os
import tempfile

import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import general, ops, yolo_head
from workloads import standin, synth
from face_detection_multi_scale_b200.datasets import letterbox
from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector
from face_detection_multi_scale_b200.preprocess_yolo_predict import preprocess_api_approach, scale_coords_api_approach
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------ letterbox (utils/datasets.py:873)
@pytest.mark.parametrize("i", range(len(H.LETTERBOX_SMALL)))
def test_letterbox_signature_and_values(golden, i):
    g = golden("letterbox")
    h, w, ns, auto, scaleup, stride, seed = H.LETTERBOX_SMALL[i]
    img = synth.synth_image(h, w, seed)
    out, ratio, pad = letterbox(img, ns, auto=auto, scaleup=scaleup, stride=stride)
    assert isinstance(out, np.ndarray) and out.dtype == np.uint8
    assert np.array_equal(out, g["small%d_out" % i])
    meta = g["small_meta"][i]
    assert ratio == (meta[8], meta[8]) and pad == (meta[9], meta[10])


def test_letterbox_scalefill_and_color_and_errors():
    img = synth.synth_image(90, 160, 3)
    out, ratio, pad = letterbox(img, (64, 96), auto=False, scaleFill=True)
    ref, rr, rp = O.letterbox(img, (64, 96), auto=False, scaleFill=True)
    assert np.array_equal(out, ref) and ratio == rr and pad == rp
    out, _, _ = letterbox(img, 128, color=(0, 0, 0), auto=False)
    ref, _, _ = O.letterbox(img, 128, color=(0, 0, 0), auto=False)
    assert np.array_equal(out, ref)
    with pytest.raises(ValueError):
        letterbox(img, 128, color=(1, 2, 300))
    with pytest.raises(ValueError):
        letterbox(img.astype(np.float32), 128)


def test_letterbox_per_channel_colour(golden):
    """cv2.copyMakeBorder takes one border value per channel (utils/datasets.py:902): reference outputs, host and
    device-resident inputs."""
    g = golden("letterbox_color")
    for i, (h, w, ns0, ns1, auto, c0, c1, c2, seed) in enumerate(g["meta"]):
        img = synth.synth_image(int(h), int(w), int(seed))
        ns = int(ns0) if ns0 == ns1 else (int(ns0), int(ns1))
        out, ratio, pad = letterbox(img, ns, color=(int(c0), int(c1), int(c2)), auto=bool(auto))
        assert np.array_equal(out, g["c%d_out" % i])
        assert (ratio[0], ratio[1], pad[0], pad[1]) == tuple(g["c%d_ratio_pad" % i])
        dev, _, _ = letterbox(torch.from_numpy(img).to(DEV), ns, color=(int(c0), int(c1), int(c2)), auto=bool(auto))
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), g["c%d_out" % i])


@pytest.mark.parametrize("i", range(len(H.API_CASES)))
def test_preprocess_api_approach_from_file(golden, i):
    cv2 = pytest.importorskip("cv2")
    g = golden("letterbox")
    h, w, size, seed = H.API_CASES[i]
    rgb = synth.synth_image(h, w, seed)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "a.png")
        cv2.imwrite(p, rgb[:, :, ::-1])
        chw = preprocess_api_approach(p, size, 32)
    assert chw.shape == (3, size, size) and chw.dtype == np.uint8
    assert sha(chw) == str(g["api%d_sha" % i])


# ------------------------------------------------------------------ head (models/yolo.py:253-308)
def test_patched_head_returns_reference_tuple(golden):
    g = golden("head")
    levels = [torch.from_numpy(g["tiny_raw%d" % i]).to(DEV) for i in range(3)]
    model = standin.SyntheticHeadModel(strides=[float(s) for s in g["tiny_strides"]], anchors_px=g["tiny_anchors_px"])
    model.to(DEV).eval()
    head = yolo_head.patch_head(yolo_head.find_head(model))
    pred, xs = head([l.clone() for l in levels])
    assert pred.shape == g["tiny_pred"].shape
    assert np.max(np.abs(pred.cpu().numpy() - g["tiny_pred"])) <= 1e-4
    for x, l in zip(xs, levels):  # (bs, na, ny, nx, no) raw maps, like the reference's second output
        bs, _, ny, nx = l.shape
        assert x.shape == (bs, 3, ny, nx, 21)
        assert torch.equal(x.permute(0, 1, 4, 2, 3).reshape(bs, 63, ny, nx), l)


# ------------------------------------------------------------------ non_max_suppression (utils/general.py:506)
def test_nms_reference_pred_bit_exact(golden):
    g = golden("head")
    pred = torch.from_numpy(g["tiny_pred"]).to(DEV)
    thr = float(g["tiny_conf_thres"])
    before = pred.clone()
    for iou in (0.45, 0.3):
        res = general.non_max_suppression(pred, thr, iou, iou_mode=ops.IOU_TV_CPU)
        assert len(res) == 2
        for b in range(2):
            ref = g["tiny_nms_iou%d_b%d" % (int(iou * 100), b)]
            assert res[b].is_cuda and res[b].shape == ref.shape
            assert np.array_equal(res[b].cpu().numpy(), ref)
    assert torch.equal(pred, before)  # does not mutate prediction


@pytest.mark.parametrize("name", list(H.CROWD_HEADS))
def test_nms_crowd_golden_through_decode(golden, name):
    g = golden("head")
    in_h, in_w, n_faces, seed = H.CROWD_HEADS[name]
    levels = [torch.from_numpy(l).to(DEV) for l in H.crowd_levels(in_h, in_w, n_faces, seed)]
    pred = yolo_head.decode_head(levels, synth.P5_STRIDES, synth.P5_ANCHORS)
    for conf, iou in ((0.5, 0.5), (0.6, 0.3), (0.25, 0.45)):
        res = general.non_max_suppression(pred, conf, iou, iou_mode=ops.IOU_TV_CPU)
        for b in range(2):
            ref = g["%s_nms_c%d_i%d_b%d" % (name, int(conf * 100), int(iou * 100), b)]
            assert res[b].shape == ref.shape
            if ref.size:
                assert np.max(np.abs(res[b].cpu().numpy() - ref)) <= 1e-4


def test_nms_empty_cpu_input_and_errors():
    pred = torch.zeros((2, 100, 21))
    res = general.non_max_suppression(pred, 0.5, 0.5)
    assert len(res) == 2 and all(r.shape == (0, 6) and not r.is_cuda for r in res)  # (0,6), general.py:527
    p = torch.from_numpy(synth.synth_pred(3000, 20, 40, 4)[None])
    r_cpu = general.non_max_suppression(p, 0.3, 0.45)[0]
    r_gpu = general.non_max_suppression(p.to(DEV), 0.3, 0.45)[0]
    assert not r_cpu.is_cuda and torch.equal(r_cpu, r_gpu.cpu()) and r_cpu.shape[0] > 5
    assert general.non_max_suppression(p.to(DEV), 0.3, 0.45, classes=[1])[0].shape == (0, 6)
    assert general.non_max_suppression(p.to(DEV), 0.3, 0.45, classes=[0])[0].shape == r_gpu.shape
    # an 80-class head without keypoints is taken now (all-zero confidences: nothing passes)
    assert general.non_max_suppression(torch.zeros((1, 10, 85), device=DEV), kpt_label=0)[0].shape == (0, 6)


def test_nms_capacity_regrow_and_max_nms():
    """More candidates than the first-try capacity (4096) and than max_nms (30000)."""
    n = 33000
    p = synth.synth_pred(n, 500, 80, 13, extent=(2000.0, 2000.0))
    p[:, 4] = np.maximum(p[:, 4], 0.9)
    p[:, 5] = np.maximum(p[:, 5], 0.9)
    res = general.non_max_suppression(torch.from_numpy(p[None]).to(DEV), 0.5, 0.5, iou_mode=ops.IOU_TV_CPU)[0]
    ref = O.non_max_suppression(p[None], 0.5, 0.5, O.TV_CPU)[0]
    assert res.shape == ref.shape and np.array_equal(res.cpu().numpy(), ref)


# ------------------------------------------------------------------ scale_coords (general.py:374, ppyp:122)
def test_scale_coords_views_in_place(golden):
    g = golden("rescale")
    ih, iw, h0, w0 = g["meta"][0]
    det = torch.zeros((37, 21), device=DEV)
    det[:, :4] = torch.from_numpy(g["c0_boxes"]).to(DEV)
    det[:, 6:] = torch.from_numpy(g["c0_kpts"]).to(DEV)
    r = general.scale_coords((ih, iw), det[:, :4], (h0, w0, 3))  # detect.py:100
    assert r.data_ptr() == det.data_ptr()
    general.scale_coords((ih, iw), det[:, 6:], (h0, w0, 3), kpt_label=5, step=3)  # detect.py:101
    assert np.array_equal(det[:, :4].cpu().numpy(), g["c0_std_boxes"])
    assert np.array_equal(det[:, 6:].cpu().numpy(), g["c0_std_kpts"])
    cpu = torch.from_numpy(g["c0_boxes"].copy())
    out = scale_coords_api_approach((ih, iw), cpu, (h0, w0, 3))
    assert out is cpu and np.array_equal(cpu.numpy(), g["c0_api_boxes"])
    empty = torch.zeros((0, 4), device=DEV)
    assert general.scale_coords((ih, iw), empty, (h0, w0)).shape == (0, 4)


# ------------------------------------------------------------------ MultiScaleFaceDetector (msfd:28-288)
def _detector(case, iou_mode=ops.IOU_TV_CPU, **kw):
    name, h, w, sizes, n_faces, conf, iou, api, seed = case
    heads, shapes = H.e2e_heads(h, w, sizes, n_faces, api, seed)
    model = standin.SyntheticHeadModel()
    for s in shapes:
        model.set_heads(s, [l[None] for l in heads[s]], pin=False)
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, conf_thres=conf, iou_thres=iou,
                                 use_api_preprocess=api, half=False, iou_mode=iou_mode, **kw)
    return det, shapes


@pytest.mark.parametrize("case", H.E2E_CASES, ids=[c[0] for c in H.E2E_CASES])
def test_detector_matches_reference_golden(golden, case):
    cv2 = pytest.importorskip("cv2")
    g = golden("e2e")
    name, h, w, sizes, n_faces, conf, iou, api, seed = case
    det, shapes = _detector(case)
    bgr = synth.synth_image(h, w, seed)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, name + ".png")
        cv2.imwrite(p, bgr)
        # model inputs (fp32, as the golden CPU run): digest of the reference's tensors
        for si, size in enumerate(sizes):
            x, img0, shape0 = det.preprocess_image(p, size)
            assert tuple(x.shape[2:]) == tuple(shapes[si]) and shape0 == (h, w, 3)
            assert sha(x.cpu().numpy()) == str(g["%s_input%d_sha" % (name, si)])
            d, shape0, dt = det.detect_single_scale(p, size)
            ref = g["%s_scale%d" % (name, si)]
            assert d.dtype == np.float64 and d.shape == ref.shape
            assert np.array_equal(d, ref)
        final, shape0 = det.detect_multi_scale(p)
    ref = g[name + "_final"]
    assert tuple(shape0) == tuple(g[name + "_shape0"])
    assert final.shape == ref.shape  # np.array([]) when nothing was found (msfd:285)
    if ref.size:
        assert np.array_equal(final, ref)
    with pytest.raises(ValueError):
        det.detect_multi_scale("/nonexistent/image.jpg")  # msfd:86


def test_detector_weighted_nms_and_weights(golden):
    g = golden("merge")
    for ci in range(5):
        n_scales, iou = g["m%d_meta" % ci]
        det = MultiScaleFaceDetector(standin.SyntheticHeadModel(), device=DEV, img_sizes=[640] * int(n_scales),
                                     iou_thres=float(iou))
        dets = g["m%d_dets" % ci]
        assert np.array_equal(det.calculate_scale_weights(dets), g["m%d_weights" % ci])
        assert np.array_equal(det.weighted_nms(dets), g["m%d_out" % ci])
    assert det.weighted_nms(np.zeros((0, 7))).shape == (0, 7)


def test_detector_batch_equals_single_and_capacity_growth():
    """B images in one fused pass == each image alone; a tiny initial capacity regrows transparently."""
    h, w, sizes, B = 270, 480, [160, 256, 320], 3
    images, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=25, seed=3, dtype=np.float16)
    model = standin.SyntheticHeadModel()
    for s, hw in enumerate(shapes):
        model.set_heads(hw, heads[s], pin=True)
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, cap=32, with_landmarks=True,
                                 iou_mode=ops.IOU_TV_CPU)
    out = det.detect_multi_scale_batch(images)
    assert det.cap > 32 and len(out) == B
    for b in range(B):
        ref, _ = O.detect_from_heads([[l[b:b + 1] for l in heads[s]] for s in range(len(sizes))], shapes, (h, w, 3),
                                     synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5, O.TV_CPU)
        assert out[b].shape == ref.shape and np.array_equal(out[b], ref)
        assert det.last_landmarks[b].shape == (len(ref), 15)
    with pytest.raises(RuntimeError):
        MultiScaleFaceDetector(model, device="cpu")


def test_nms_other_keypoint_counts(golden):
    """kpt_label = 0 (no = 6: the head without landmarks / the [..., :6] call of msfd:134) and kpt_label = 2."""
    g, h = golden("nms6"), golden("head")
    thr = float(g["thr"])
    for cols, kl in ((6, 0), (12, 2)):
        pred = torch.from_numpy(np.ascontiguousarray(h["tiny_pred"][..., :cols])).to(DEV)
        res = general.non_max_suppression(pred, thr, 0.45, kpt_label=kl, iou_mode=ops.IOU_TV_CPU)
        for b in range(2):
            assert np.array_equal(res[b].cpu().numpy(), g["c%d_b%d" % (cols, b)])
    # kpt_label=5 on 6 columns (msfd:134) gives nc = -14 in the reference: its confidence is then the RAW class score of
    # the rows whose objectness passes (pinned against the reference's function on the GPU box: tests/ref_gpu_check.py P2b)
    p6 = torch.from_numpy(np.ascontiguousarray(h["tiny_pred"][..., :6])).to(DEV)
    res = general.non_max_suppression(p6, thr, 0.45, iou_mode=ops.IOU_TV_CPU)
    for b in range(2):
        x = h["tiny_pred"][b][:, :6]
        x = x[(x[:, 4] > np.float32(thr)) & (x[:, 5] > np.float32(thr))]
        r = res[b].cpu().numpy()
        assert len(r) <= len(x) and (len(x) == 0 or len(r) > 0)
        assert np.all(np.isin(r[:, 4], x[:, 5]))  # conf column = one of the raw class scores
    with pytest.raises(NotImplementedError):  # several classes AND keypoints: the reference itself only reads class 0 there
        general.non_max_suppression(torch.zeros((1, 10, 5 + 3 + 15), device=DEV), kpt_label=5)
    # multi-class heads without keypoints (kpt_label=0, no = 5 + nc): against a direct evaluation with torchvision
    import torchvision
    g_ = torch.Generator().manual_seed(1)
    p = torch.cat((torch.rand(1, 500, 2, generator=g_) * 300, torch.rand(1, 500, 2, generator=g_) * 50 + 10, torch.rand(1, 500, 1, generator=g_),
                   torch.rand(1, 500, 3, generator=g_)), 2).to(DEV)
    out = general.non_max_suppression(p, 0.25, 0.5, kpt_label=0)[0]
    x = p[0][p[0, :, 4] > 0.25].clone()
    x[:, 5:] *= x[:, 4:5]
    conf, j = x[:, 5:].max(1, keepdim=True)
    x = torch.cat((general.xywh2xyxy(x[:, :4]), conf, j.float()), 1)[conf.view(-1) > 0.25]
    keep = torchvision.ops.nms(x[:, :4] + x[:, 5:6] * 4096, x[:, 4], 0.5)
    assert torch.equal(out, x[keep]) and len(out) > 20 and len(torch.unique(out[:, 5])) == 3


def test_detect_stream_matches_batch_calls():
    """Pipelined multi-batch API == one synchronous call per batch (different content per batch, 2 and 3 lanes)."""
    h, w, sizes, B = 270, 480, [160, 256, 320], 2
    images, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=20, seed=9, dtype=np.float16)
    model = standin.SyntheticHeadModel()
    for s, hw in enumerate(shapes):
        model.set_heads(hw, heads[s], pin=True)
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU)
    batches = [np.roll(images, k, axis=2) if k else images for k in range(5)]
    ref = [det.detect_multi_scale_batch(b) for b in batches]
    for depth in (1, 2, 3):
        got = list(det.detect_stream([torch.from_numpy(b).pin_memory() for b in batches], depth=depth))
        assert len(got) == len(ref)
        for g_, r_ in zip(got, ref):
            for a, b in zip(g_, r_):
                assert a.shape == b.shape and np.array_equal(a, b)
    det2 = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU, cap=16)  # overflow path
    got = list(det2.detect_stream(batches[:2]))
    for g_, r_ in zip(got, ref[:2]):
        for a, b in zip(g_, r_):
            assert np.array_equal(a, b)


def test_detector_shared_by_thread_pool():
    """One detector served from a ThreadPoolExecutor (the reference's directory mode,
    yolov7_face_multi_scale_dataframe_predict.py:279-282): per-thread native contexts, plan caches and head mode."""
    import concurrent.futures
    h, w, sizes = 270, 480, [160, 256, 320]
    images, heads, shapes = standin.make_batch(h, w, sizes, 1, n_faces=20, seed=12, dtype=np.float16)
    model = standin.SyntheticHeadModel()
    for s, hw in enumerate(shapes):
        model.set_heads(hw, heads[s], pin=False, device=DEV)
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU)
    ref, _ = det.detect_multi_scale(images[0])

    def work(i):
        out, shape0 = det.detect_multi_scale(images[0])
        pred, xs = det.head([l.clone() for l in model._heads[tuple(shapes[i % 3])]])  # compat-level call in between
        return out, pred.shape

    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(work, range(24)))
    for out, pshape in results:
        assert out.shape == ref.shape and np.array_equal(out, ref)
        assert pshape[2] == 21


def test_non_max_suppression_with_apriori_labels(golden):
    """general.non_max_suppression(labels=...) (utils/general.py:533-541, autolabelling) against the reference's own
    output; with keypoint columns the reference's torch.cat raises, here a ValueError."""
    g, h = golden("labels"), golden("head")
    pred = torch.from_numpy(np.ascontiguousarray(h["tiny_pred"][..., :6])).to(DEV)
    labs = [torch.from_numpy(g["l0"]).to(DEV), torch.from_numpy(g["l1"]).to(DEV)]
    res = general.non_max_suppression(pred, float(g["thr"]), 0.45, kpt_label=0, labels=labs, iou_mode=ops.IOU_TV_CPU)
    for b in range(2):
        assert np.array_equal(res[b].cpu().numpy(), g["r%d" % b])
    with pytest.raises(ValueError):
        general.non_max_suppression(torch.from_numpy(h["tiny_pred"]).to(DEV), 0.3, 0.45, labels=labs)


def test_load_images_matches_reference_sequence(tmp_path):
    """datasets.LoadImages (utils/datasets.py:124-201): letterbox(auto=False) + BGR->RGB + CHW of every file, against
    the oracle's restatement of the same cv2 / numpy sequence on the decoded file (PNG: lossless)."""
    import cv2
    from face_detection_multi_scale_b200.datasets import LoadImages
    from oracle import oracle as O
    shapes = [(120, 160), (200, 90), (64, 64)]
    for i, (h, w) in enumerate(shapes):
        assert cv2.imwrite(str(tmp_path / ("im%d.png" % i)), synth.synth_image(h, w, 40 + i))
    (tmp_path / "notes.txt").write_text("not an image")
    ds = LoadImages(str(tmp_path), img_size=96, stride=32)
    assert len(ds) == 3 and ds.mode == "image"
    seen = 0
    for path, img, img0, cap in ds:
        assert cap is None and img0.shape[:2] == shapes[seen]
        ref, _, _ = O.letterbox(img0, 96, auto=False, stride=32)
        assert np.array_equal(img, np.ascontiguousarray(ref[:, :, ::-1].transpose(2, 0, 1)))
        seen += 1
    assert seen == 3
    (p, dev_img, _, _) = next(iter(LoadImages(str(tmp_path / "im0.png"), img_size=96, device_out=True)))
    assert dev_img.is_cuda and dev_img.shape == (3, 96, 96)


def test_read_batch_and_detect_from_paths(tmp_path):
    """datasets.read_batch: files decoded in parallel into one pinned batch, identical to np.stack(cv2.imread(..));
    detect_multi_scale_batch(list of paths) goes through it."""
    import cv2
    from face_detection_multi_scale_b200.datasets import read_batch
    paths = []
    for i in range(5):
        p = tmp_path / ("im%d.png" % i)
        assert cv2.imwrite(str(p), synth.synth_image(96, 160, 70 + i))
        paths.append(str(p))
    batch = read_batch(paths, threads=3)
    assert batch.is_pinned() and tuple(batch.shape) == (5, 96, 160, 3)
    assert np.array_equal(batch.numpy(), np.stack([cv2.imread(p) for p in paths]))
    again = read_batch(paths[::-1], out=batch)
    assert again.data_ptr() == batch.data_ptr() and np.array_equal(again.numpy()[0], cv2.imread(paths[-1]))
    with pytest.raises(ValueError):
        read_batch([paths[0], synth.synth_image(64, 64, 1)])
    with pytest.raises(ValueError):
        read_batch([str(tmp_path / "missing.png")])


def test_detect_files_pipelined_equals_per_file(tmp_path):
    """detect_files (parallel host decode -> pinned batches -> detect_stream) yields, in order, exactly what
    detect_multi_scale returns file by file; the last batch is ragged."""
    import cv2
    h, w, sizes = 192, 320, [160, 320]
    n = 7
    images, heads, shapes = standin.make_batch(h, w, sizes, n, n_faces=6, seed=3)
    paths = []
    for i in range(n):
        p = tmp_path / ("f%02d.png" % i)
        assert cv2.imwrite(str(p), images[i])
        paths.append(str(p))

    def detector(batch):
        model = standin.SyntheticHeadModel()
        for s_, hw in enumerate(shapes):
            model.set_heads(hw, [l[batch] for l in heads[s_]], device=DEV)
        return MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, conf_thres=0.4, iou_thres=0.5, cap=512)

    # the stand-in backbone serves fixed head maps per batch position: run files in batches of the same composition
    got = {}
    for lo in range(0, n, 3):
        idx = list(range(lo, min(lo + 3, n)))
        det = detector(idx)
        for pth, d in det.detect_files([paths[i] for i in idx], batch_size=3, decode_threads=2):
            got[pth] = d
    assert list(got) == paths
    for i, pth in enumerate(paths):
        one = detector([i])
        ref, shape0 = one.detect_multi_scale(pth)
        assert shape0 == (h, w, 3)
        assert np.array_equal(np.asarray(got[pth]), np.asarray(ref))


def test_compat_ops_fuzz_vs_oracle():
    """Randomised run of the call-surface functions against the oracle: non_max_suppression (no = 21 / 6, fp16 / fp32
    predictions, CPU and CUDA tensors, classes filter), scale_coords (boxes / keypoints, with and without ratio_pad),
    scale_coords_api_approach, scale_img / descale_pred, bbox_overlaps."""
    from face_detection_multi_scale_b200 import torch_utils as TU
    from face_detection_multi_scale_b200.evaluation import bbox_overlaps
    rng = np.random.default_rng(4242)
    for it in range(16):
        ops.debug_poison_smem(0xFF if it & 1 else 0x7F)
        # --- non_max_suppression
        bs, rows = int(rng.integers(1, 4)), int(rng.choice([1, 7, 300, 2500, 6000]))
        no = 21 if it % 3 else 6
        pred = np.stack([synth.synth_pred(rows, int(rng.integers(1, 12)), int(rng.integers(1, 40)), 100 * it + b)
                         for b in range(bs)])
        pred = np.ascontiguousarray(pred[..., :no])
        conf, iou = float(rng.choice([0.1, 0.3, 0.5])), float(rng.choice([0.3, 0.45, 0.6]))
        half = it % 4 >= 2  # fp16 predictions: up-cast, filtered / converted / suppressed in fp32, kept rows rounded back
        if half:
            pred = pred.astype(np.float16)
        t = torch.from_numpy(pred)
        if it & 1:
            t = t.to(DEV)
        got = general.non_max_suppression(t, conf, iou, kpt_label=5 if no == 21 else 0, nc=1,
                                          iou_mode=ops.IOU_TV_CPU)
        ref = O.non_max_suppression(pred.astype(np.float32), conf, iou, O.TV_CPU)
        assert len(got) == bs
        for b in range(bs):
            assert got[b].is_cuda == bool(it & 1)
            g_, r_ = got[b].cpu().numpy(), ref[b]
            assert g_.shape == r_.shape, (it, b, g_.shape, r_.shape)
            if half and len(r_):
                assert got[b].dtype == torch.float16
                r_ = r_.astype(np.float16)
            assert np.array_equal(g_, r_), (it, b)
        # --- scale_coords / scale_coords_api_approach
        ih, iw = int(rng.integers(2, 40)) * 32, int(rng.integers(2, 40)) * 32
        h0, w0 = int(rng.integers(30, 3000)), int(rng.integers(30, 3000))
        n = int(rng.choice([0, 1, 3, 4, 5, 200]))
        boxes = rng.uniform(-50, max(ih, iw) + 50, (n, 4)).astype(np.float32)
        kpts = rng.uniform(-50, max(ih, iw) + 50, (n, 15)).astype(np.float32)
        rp = None if it % 2 else ((float(rng.uniform(0.1, 2.0)),) * 2, (float(rng.uniform(0, 40)), float(rng.uniform(0, 40))))
        gb = general.scale_coords((ih, iw), torch.from_numpy(boxes.copy()).to(DEV), (h0, w0), rp)
        assert np.array_equal(gb.cpu().numpy(), O.scale_coords((ih, iw), boxes.copy(), (h0, w0), rp))
        gk = general.scale_coords((ih, iw), torch.from_numpy(kpts.copy()), (h0, w0), rp, kpt_label=5, step=3)
        assert np.array_equal(gk.numpy(), O.scale_coords((ih, iw), kpts.copy(), (h0, w0), rp, kpt_label=5, step=3))
        ga = scale_coords_api_approach((ih, iw), torch.from_numpy(boxes.copy()).to(DEV), (h0, w0, 3))
        assert np.array_equal(ga.cpu().numpy(), O.scale_coords_api((ih, iw), boxes.copy(), (h0, w0)))
        # --- test-time augmentation pieces
        shape = (int(rng.integers(1, 3)), 3, int(rng.integers(1, 8)) * 32, int(rng.integers(1, 8)) * 32)
        x = rng.random(shape, dtype=np.float32)
        ratio, flip = float(rng.choice([1.0, 0.83, 0.67, 0.5])), [None, 2, 3][int(rng.integers(0, 3))]
        gs_img = TU.scale_img(torch.from_numpy(x).to(DEV), ratio, gs=32, flip=flip).cpu().numpy()
        assert np.array_equal(gs_img, O.scale_img(x, ratio, gs=32, flip=flip)), (it, shape, ratio, flip)
        p = rng.uniform(0, 300, (shape[0], int(rng.integers(1, 500)), 21)).astype(np.float32)
        gd = TU.descale_pred(torch.from_numpy(p.copy()).to(DEV), flip, ratio, shape[2:]).cpu().numpy()
        assert np.array_equal(gd, O.descale_pred(p.copy(), flip, ratio, shape[2:])), (it, "descale")
        # --- evaluation IoU matrix (float64, +1 convention)
        a = rng.uniform(0, 800, (int(rng.integers(1, 400)), 4)); a[:, 2:] += a[:, :2]
        q = rng.uniform(0, 800, (int(rng.integers(1, 90)), 4)); q[:, 2:] += q[:, :2]
        if it % 4 == 0:
            a, q = np.round(a), np.round(q)
        assert np.array_equal(bbox_overlaps(a, q), O.bbox_overlaps(a, q))


def test_c_program_drives_the_library_on_the_gpu(tmp_path):
    """include/fdms.h from plain C on the GPU box: create -> fdms_nms on buffers the program allocates itself with the
    CUDA runtime -> destroy; the keep list equals torchvision's order (highest score first, ties / suppression as a7)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if cc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("no C toolchain / CUDA headers on this box")
    exe = str(tmp_path / "c_abi_gpu")
    lib_dir = os.path.join(root, "face_detection_multi_scale_b200")
    subprocess.check_call([cc, "-std=c11", "-Wall", "-o", exe, os.path.join(root, "tests", "c_abi_gpu.c"), "-I", os.path.join(root, "include"),
                           "-I", os.path.join(cuda, "include"), "-L", lib_dir, "-l:libfdms.so", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "n_keep 3 keep 4 0 5", lines  # 0.95 (suppresses 3), 0.9 (suppresses 2 and 1), 0.3
    assert lines[1] == "letterbox 0 384 640 top 12", lines
    assert lines[2].startswith("null rows -> -1"), lines


def test_detect_stream_landmarks_and_capacity_growth():
    """Streaming path: `last_landmarks` belongs to the batch that was yielded last (not to an earlier synchronous call),
    and after a batch overflowed `out_cap` the lanes rebuild their plans instead of overflowing for ever."""
    h, w, sizes, B = 270, 480, [160, 256, 320], 2
    images, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=25, seed=14, dtype=np.float16)
    model = standin.SyntheticHeadModel()
    for s, hw in enumerate(shapes):
        model.set_heads(hw, heads[s], pin=True)
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU, with_landmarks=True)
    batches = [images, np.ascontiguousarray(images[::-1])]
    want = []
    for b in batches:  # the heads do not depend on the pixels: batch 1 is batch 0 with the images swapped
        det.detect_multi_scale_batch(images)
        want.append(([k.copy() for k in det.last_landmarks], dict(det.last_stats)))
    det.last_landmarks = "stale"
    for i, dets in enumerate(det.detect_stream(batches, depth=2)):
        assert isinstance(det.last_landmarks, list) and len(det.last_landmarks) == B
        for b in range(B):
            assert det.last_landmarks[b].shape == (len(dets[b]), 15)
            assert np.array_equal(det.last_landmarks[b], want[i][0][b])
        assert det.last_stats["final"] == want[i][1]["final"]
    # out_cap too small for the merged rows: the first streamed batch is redone synchronously and grows it ...
    det2 = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU, cap=256, out_cap=8)
    ref = det.detect_multi_scale_batch(images)
    got = list(det2.detect_stream([images] * 6, depth=2))
    for g_ in got:
        for a, b in zip(g_, ref):
            assert np.array_equal(a, b)
    assert det2.out_cap > 8
    # ... and the lanes' plans were rebuilt with the grown capacity (no plan of the old size is left)
    for lane in det2._tls.lanes:
        assert lane.plans and all(p.out.out_cap == det2.out_cap for p in lane.plans.values())


def _write_video(path, frames):
    import cv2
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), 25, (w, h))
    if not vw.isOpened():
        pytest.skip("this OpenCV build cannot write MJPG/avi")
    for f in frames:
        vw.write(f)
    vw.release()


def test_load_images_video_frames_and_mixed_directory(tmp_path):
    """utils/datasets.py:160-178: every frame of every clip after the stills, `mode` / `frame` / `nframes` / `count` as
    the reference reports them; pixels = the oracle's letterbox(auto=False) sequence on the decoded frame."""
    import cv2
    from face_detection_multi_scale_b200.datasets import LoadImages
    from oracle import oracle as O
    frames = [np.full((90, 120, 3), 30 * i, np.uint8) + synth.synth_image(90, 120, i) // 4 for i in range(5)]
    _write_video(tmp_path / "b.avi", frames)
    _write_video(tmp_path / "c.avi", frames[:2])
    assert cv2.imwrite(str(tmp_path / "a.png"), synth.synth_image(64, 100, 9))
    ds = LoadImages(str(tmp_path), img_size=64, stride=32)
    assert ds.nf == 3 and ds.video_flag == [False, True, True] and ds.cap is not None
    seen = []
    for path, img, img0, cap in ds:
        ref, _, _ = O.letterbox(img0, 64, auto=False, stride=32)
        assert np.array_equal(img, np.ascontiguousarray(ref[:, :, ::-1].transpose(2, 0, 1)))
        seen.append((os.path.basename(path), ds.mode, ds.frame if ds.mode == "video" else 0))
        if ds.mode == "video":
            assert cap is ds.cap and ds.nframes in (5, 2)
    assert seen == [("a.png", "image", 0)] + [("b.avi", "video", k) for k in range(1, 6)] + [("c.avi", "video", 1), ("c.avi", "video", 2)]
    assert ds.count == ds.nf
    # decoded frames equal what cv2 gives when the file is read directly
    cap = cv2.VideoCapture(str(tmp_path / "b.avi"))
    ok, f0 = cap.read()
    first = next(iter(LoadImages(str(tmp_path / "b.avi"), img_size=64)))
    assert ok and np.array_equal(first[2], f0)


def test_load_streams_batches_frames_like_the_reference(tmp_path):
    """utils/datasets.py:262-340: the stream loader (the reference's only batch > 1 loader).  File sources stand in for
    cameras; the first tick holds the first frame of every source (`cap.read()` at :294), letterboxed with auto = rect,
    RGB, CHW, stacked."""
    from face_detection_multi_scale_b200.datasets import LoadStreams, clean_str
    from oracle import oracle as O
    fa = [synth.synth_image(90, 120, 20 + i) for i in range(40)]
    fb = [synth.synth_image(90, 120, 70 + i) for i in range(40)]
    _write_video(tmp_path / "cam_a.avi", fa)
    _write_video(tmp_path / "cam_b.avi", fb)
    lst = tmp_path / "streams.txt"
    lst.write_text("%s\n\n%s\n" % (tmp_path / "cam_a.avi", tmp_path / "cam_b.avi"))
    ds = LoadStreams(str(lst), img_size=64, stride=32)
    assert ds.mode == "stream" and ds.rect and len(ds) == 0
    assert ds.sources == [clean_str(str(tmp_path / "cam_a.avi")), clean_str(str(tmp_path / "cam_b.avi"))]
    srcs, img, img0, cap = next(iter(ds))
    assert cap is None and img.shape[0] == 2 and img.dtype == np.uint8 and len(img0) == 2
    for i in range(2):
        ref, _, _ = O.letterbox(img0[i], 64, auto=True, stride=32)
        assert np.array_equal(img[i], np.ascontiguousarray(ref[:, :, ::-1].transpose(2, 0, 1)))
    # a later tick still yields one frame per source with the same shape; on the device when asked for
    ds2 = LoadStreams(str(tmp_path / "cam_a.avi"), img_size=64, device_out=True)
    it = iter(ds2)
    _, img_a, _, _ = next(it)
    _, img_b, _, _ = next(it)
    assert img_a.is_cuda and tuple(img_a.shape) == tuple(img_b.shape) == (1, 3, 64, 64) or tuple(img_a.shape)[0] == 1


def test_ragged_batch_mixed_image_sizes(tmp_path):
    """Images of different sizes in ONE call (msfd:242-288 takes any image; autoShape batches mixed sizes,
    models/common.py:618-636): bucketed by letterboxed shapes, per-image original size in the merge.  Every image's
    result equals the oracle's for that image alone; detect_files takes a directory of mixed sizes too."""
    import cv2
    sizes = [160, 256, 320]
    src_shapes = [(270, 480), (260, 480), (480, 270), (270, 480), (300, 480)]  # (270,480) and (260,480) share a bucket
    imgs = [synth.synth_image(h, w, 60 + i) for i, (h, w) in enumerate(src_shapes)]
    det = None
    per_img = []
    for i, (h, w) in enumerate(src_shapes):
        hd, shp = H.e2e_heads(h, w, sizes, 15, False, 900 + i)
        per_img.append((hd, shp))
    # the stand-in backbone hands out pre-generated maps per input shape: build them in the order the detector batches
    # a bucket (images sorted by source shape inside a bucket)
    buckets = {}
    for i, (hd, shp) in enumerate(per_img):
        buckets.setdefault(tuple(shp), []).append(i)
    model = standin.SyntheticHeadModel()

    class Routed(standin.SyntheticHeadModel):
        """Delivers the maps of whichever bucket matches the input's shape AND batch size."""
        def __init__(self):
            super().__init__()
            self.by_key = {}

        def forward(self, x, augment=False):
            levels = self.by_key[(int(x.shape[0]), int(x.shape[2]), int(x.shape[3]))]
            return self.model[-1]([t.to(x.device) for t in levels])
    model = Routed()
    for shp, idx in buckets.items():
        idx = sorted(idx, key=lambda i: imgs[i].shape)
        for s_, hw in enumerate(shp):
            levels = [torch.from_numpy(np.stack([per_img[i][0][hw][l] for i in idx]).astype(np.float16)) for l in range(3)]
            model.by_key[(len(idx), hw[0], hw[1])] = levels
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU)
    out = det.detect_multi_scale_batch(imgs)
    assert det.last_stats["buckets"] == len(buckets) >= 3
    total = 0
    for i, (h, w) in enumerate(src_shapes):
        hd, shp = per_img[i]
        ref, _ = O.detect_from_heads([[hd[hw][l][None].astype(np.float16) for l in range(3)] for hw in shp], shp, (h, w, 3),
                                     synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5, O.TV_CPU)
        assert out[i].shape == ref.shape and (ref.size == 0 or np.array_equal(out[i], ref)), (i, out[i].shape, ref.shape)
        total += len(ref)
    assert total > 20
    # from files, through detect_files (chunk of mixed sizes)
    paths = []
    for i, im in enumerate(imgs):
        p = tmp_path / ("r%d.png" % i)
        assert cv2.imwrite(str(p), im)
        paths.append(str(p))
    got = list(det.detect_files(paths, batch_size=8))
    assert [p for p, _ in got] == paths
    for (p, d), o in zip(got, out):
        assert d.shape == o.shape and np.array_equal(d, o)
