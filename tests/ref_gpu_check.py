#!/usr/bin/env python
"""The drop-in against the UNMODIFIED reference with a REAL reference network, on the GPU (run as a subprocess by
tests/test_gpu_reference_model.py so that the reference's top-level `utils` / `models` packages stay out of pytest).

The reference tree is the byte-for-byte copy `baseline/_ref/` (or `/root/reference` in the build container).  A
random-init `models.yolo.Model('cfg/yolov7s-face.yaml')` (seeded) is saved as a checkpoint, so the reference's own
constructor (`MultiScaleFaceDetector(model_path, ...)`, multi_scale_face_detector.py:33-58 -> load_yolo_model ->
attempt_load) and the unmodified `detect.py` run exactly as a user runs them.  Random-init heads give obj*cls <= ~0.25
(SURVEY F6), so the network's BatchNorm statistics are set from random batches (baseline/ref_runner.py:random_init_model):
about 1.5 % of the rows then pass the reference's default conf_thres 0.5.

  Everything below runs with FDMS_NUMERICS=aten_cuda (include/fdms.h): the sigmoid uses the expf ATen's CUDA kernel
  calls and a division by a Python scalar is a multiplication by its fp32 reciprocal, as on a CUDA device.
  P  stage by stage on IDENTICAL inputs, fp32:
     P1 head decode: the patched head (one kernel) vs the reference's IKeypoint.forward (ATen CUDA elementwise ops) on the
        same conv outputs -> bit-identical.  Also reported (not the default any more on a GPU, but what the CPU oracle
        pins): the same comparison with FDMS_NUMERICS=exact -- the correctly rounded exponential against libdevice's
        expf: <= 8 ulp, ulp histogram in the report
     P2 non_max_suppression on the reference's own decoded tensor -> bit-identical lists (candidate sets, keep order, rows)
     P3 scale_coords_api_approach(...).round() and scale_coords on identical rows -> bit-identical
  B  monkey patch (INTEGRATION route a), fp32: the reference's OWN class with letterbox / non_max_suppression /
     scale_coords_api_approach rebound and the head decode patched, against its unpatched run -> the same (k, 7) arrays
  A  class swap (route b), fp32: face_detection_multi_scale_b200.MultiScaleFaceDetector against the reference's class,
     same checkpoint, same files, same GPU -> the same (k, 7) arrays
     P1h the same for the reference's DEFAULT CUDA mode, a half model (msfd:49): its conv outputs, sigmoid, `* 2.`, `- 0.5`,
        `** 2`, `* anchor_grid` are half tensors, `+ grid` / `* stride` promote to float32, pred is float32 -- the kernels
        round where ATen rounds (csrc/common.cuh h_sigmoid / h_xy / h_wh) -> bit-identical pred
  T  test-time augmentation: Model.forward(x, augment=True) against torch_utils.forward_augment, fp32 and half model -> identical
  C  class swap against the reference's default CUDA mode (half precision) -> the same (k, 7) arrays
  C2 class swap with the production preprocessing (use_api_preprocess=True), fp32 and half -> the same (k, 7) arrays
  D  unmodified detect.py (half, --save-txt --save-conf): label files of the patched run vs the unpatched run -> identical
  E  timings: backbone forward per scale, reference GPU path per image, libfdms path per image

Prints one JSON line {"checks": n, "failures": [...], "report": {...}}; exit 1 on failures.
"""
import argparse
import contextlib
import glob
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
os.environ["FDMS_NUMERICS"] = "aten_cuda"  # read when a native context is created (include/fdms.h)

from baseline import refshim  # noqa: E402

FAIL, N, REPORT = [], [0], {}


def check(ok, what):
    N[0] += 1
    if not ok:
        FAIL.append(what)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def match_boxes(a, b, thr=0.9):
    """Fraction of rows of `a` that have a row of `b` with IoU >= thr (greedy, for the fp16 comparisons)."""
    if len(a) == 0 or len(b) == 0:
        return 1.0 if len(a) == len(b) else 0.0
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iou = inter / np.maximum(aa[:, None] + ab[None] - inter, 1e-9)
    return float(np.mean(iou.max(1) >= thr))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg/yolov7s-face.yaml")
    ap.add_argument("--sizes", default="640,960,1280")
    ap.add_argument("--conf", type=float, default=0.5, help="the reference's default confidence threshold (msfd:33)")
    ap.add_argument("--no-detect-py", action="store_true")
    ap.add_argument("--quick", action="store_true", help="P and A only")
    args = ap.parse_args()
    if not refshim.install():
        print(json.dumps({"checks": 0, "failures": [], "skipped": "no reference tree"}))
        return 0
    import cv2
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"checks": 0, "failures": [], "skipped": "no CUDA device"}))
        return 0
    torch.backends.cudnn.benchmark = False
    ref_root = refshim.REFERENCE_ROOT
    with contextlib.redirect_stdout(io.StringIO()):
        import multi_scale_face_detector as msfd
        from models.yolo import Model
    import face_detection_multi_scale_b200.integrate as fi
    from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector as Mine
    from workloads import synth

    sizes = [int(v) for v in args.sizes.split(",")]
    tmp = tempfile.mkdtemp(prefix="fdms_ref_")
    from baseline.ref_runner import random_init_model
    net = random_init_model(args.cfg, seed=0)  # random init + BatchNorm statistics from random batches: well conditioned
    ckpt = os.path.join(tmp, "random_init.pt")
    torch.save({"model": net}, ckpt)
    REPORT["cfg"], REPORT["params_M"] = args.cfg, round(sum(p.numel() for p in net.parameters()) / 1e6, 2)

    # inputs: the reference's own crowd photographs that travelled + one synthetic 1080p frame
    files = sorted(glob.glob(os.path.join(ref_root, "data", "images", "*.jpg")))[:3]
    syn = os.path.join(tmp, "synthetic_1080p.png")
    cv2.imwrite(syn, synth.synth_image(1080, 1920, 5))
    files.append(syn)
    REPORT["files"] = [os.path.basename(f) for f in files]

    def ref_detector(half, api=False):
        d = quiet(msfd.MultiScaleFaceDetector, ckpt, "", sizes, 0.5, 0.5, api)  # the reference's own constructor
        if not half:
            # load_yolo_model has already rounded the weights to fp16 (ppyp:395): the fp32 mode needs the checkpoint's own
            # weights, as its loader returns them
            from models.experimental import attempt_load
            d.half = False
            d.model = quiet(attempt_load, ckpt, map_location=d.device).float().eval()
        return d

    conf = args.conf
    ref32 = ref_detector(False)
    ref32.conf_thres = conf
    REPORT["conf_thres"] = conf

    def rows_equal(a, b, tag):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if a.size == 0 or b.size == 0:
            return a.size == b.size
        a, b = a.reshape(-1, 7), b.reshape(-1, 7)
        if a.shape != b.shape:
            return False
        return bool(np.array_equal(a, b))

    # ---- P: stage by stage on identical inputs, fp32
    import utils.general as ref_general
    import utils.preprocess_yolo_predict as ref_pp
    from face_detection_multi_scale_b200 import general as G, preprocess_yolo_predict as PP, yolo_head
    import copy as _copy
    model_p = _copy.deepcopy(ref32.model)
    yolo_head.patch_head(yolo_head.find_head(model_p))
    hist = np.zeros(8, np.int64)
    worst_px, worst_ulp = 0.0, 0
    mine_p = Mine(ckpt, "", sizes, conf, 0.5, half=False)
    p0 = []
    for f in (files[0], files[-1]):
        for size in sizes:
            x, _, shape0 = quiet(ref32.preprocess_image, f, size)
            # P0: the network input -- the fused pyramid (fp32, aten_cuda: x * (1/255)) against letterbox + pack + `/= 255.0`
            xm, _, _ = mine_p.preprocess_image(f, size)
            same_in = tuple(x.shape) == tuple(xm.shape) and torch.equal(x, xm)
            if not same_in and tuple(x.shape) == tuple(xm.shape):
                u8 = torch.round(xm * 255.0)
                p0.append({"file": os.path.basename(f), "size": size, "differing": int((x != xm).sum()), "max_abs": float((x - xm).abs().max()),
                           "ref_is_mul_recip": bool(torch.equal(x, u8 * (1.0 / 255.0))), "ref_is_true_div": bool(torch.equal(x, u8 / 255.0))})
            check(same_in, ("P0 network input", os.path.basename(f), size, p0[-1] if p0 and not same_in else None))
            with torch.no_grad():
                pred_ref = ref32.model(x)[0]
                pred_mine = model_p(x)[0]
            a_, b_ = pred_ref.float().cpu().numpy(), pred_mine.float().cpu().numpy()
            d = np.abs(a_ - b_)
            ulp = np.abs(a_.view(np.int32).astype(np.int64) - b_.view(np.int32).astype(np.int64))
            check(a_.shape == b_.shape and int(ulp.max()) == 0, ("P1 decode bit-identical (aten_cuda)", os.path.basename(f), size, int(ulp.max())))
            # the oracle's flavour (correctly rounded exp) against ATen's CUDA decode, for the record
            from face_detection_multi_scale_b200 import _lib as _L
            cx_ = _L.ctx(torch.cuda.current_device())
            cx_.set_numerics("exact")
            with torch.no_grad():
                pe = model_p(x)[0].float().cpu().numpy()
            cx_.set_numerics("aten_cuda")
            ue = np.abs(a_.view(np.int32).astype(np.int64) - pe.view(np.int32).astype(np.int64))
            de = np.abs(a_ - pe)
            hist += np.bincount(np.minimum(ue, 7).ravel(), minlength=8)
            worst_ulp = max(worst_ulp, int(ue.max()))
            worst_px = max(worst_px, float(de[np.abs(a_) < 256].max()))
            rel = de / np.maximum(np.abs(a_), 1e-30)
            check(float(de[np.abs(a_) < 256].max()) <= 1e-4 and float(rel[np.abs(a_) >= 256].max() if (np.abs(a_) >= 256).any() else 0.0) <= 1e-6,
                  ("P1 decode exact-mode vs ATen CUDA", os.path.basename(f), size, float(de.max())))
            # P2: both non_max_suppression implementations on the reference's own tensor
            r_list = quiet(ref_general.non_max_suppression, pred_ref.clone(), conf, 0.5)
            m_list = G.non_max_suppression(pred_ref.clone(), conf, 0.5)
            same = len(r_list) == len(m_list) and all(tuple(r.shape) == tuple(m.shape) and torch.equal(r, m) for r, m in zip(r_list, m_list))
            check(same, ("P2 nms on identical input", os.path.basename(f), size, [tuple(r.shape) for r in r_list], [tuple(m.shape) for m in m_list]))
            # P2b: the `pred[..., :6]` fallback of msfd:134 (default kpt_label=5 on six columns: conf = raw class score)
            r6 = quiet(ref_general.non_max_suppression, pred_ref[..., :6].clone(), conf, 0.5)
            m6 = G.non_max_suppression(pred_ref[..., :6].clone(), conf, 0.5)
            check(len(r6) == len(m6) and all(tuple(r.shape) == tuple(m.shape) and torch.equal(r, m) for r, m in zip(r6, m6)),
                  ("P2b nms on pred[..., :6]", os.path.basename(f), size, [tuple(r.shape) for r in r6], [tuple(m.shape) for m in m6]))
            # P3: rescale on identical rows
            if len(r_list[0]):
                r_box = quiet(ref_pp.scale_coords_api_approach, x.shape[2:], r_list[0][:, :4].clone(), shape0).round()
                m_box = PP.scale_coords_api_approach(x.shape[2:], r_list[0][:, :4].clone(), shape0).round()
                r_sc = ref_general.scale_coords(x.shape[2:], r_list[0][:, :4].clone(), shape0)
                m_sc = G.scale_coords(x.shape[2:], r_list[0][:, :4].clone(), shape0)
                check(torch.equal(r_box, m_box), ("P3 scale_coords_api_approach", os.path.basename(f), size))
                check(torch.equal(r_sc, m_sc), ("P3 scale_coords", os.path.basename(f), size, float((r_sc - m_sc).abs().max())))
    # T: test-time augmentation (models/yolo.py:363-374: scale_img -> forward_once -> _descale_pred, three passes, cat) around
    # the same network: the reference's Model.forward(x, augment=True) against torch_utils.forward_augment on libfdms
    # (fdms_scale_img with ATen's CUDA contraction, the patched head, fdms_descale_pred)
    from face_detection_multi_scale_b200 import torch_utils as TU
    t_rep = []
    for f in (files[0], files[-1]):
        x, _, _ = quiet(ref32.preprocess_image, f, sizes[0])
        with torch.no_grad():
            tr = ref32.model(x, augment=True)[0]
            tm = TU.forward_augment(model_p.forward_once, x, gs=int(ref32.model.stride.max()))[0]
        same = tuple(tr.shape) == tuple(tm.shape) and torch.equal(tr, tm)
        t_rep.append({"file": os.path.basename(f), "rows": int(tr.shape[1]), "bit_identical": bool(same),
                      "max_abs": float((tr - tm).abs().max()) if tuple(tr.shape) == tuple(tm.shape) else None})
        check(same, ("T forward_augment fp32", t_rep[-1]))
    REPORT["T_forward_augment_fp32"] = t_rep
    # P1h: the head decode of a HALF model on identical half conv outputs (the reference's default mode on a GPU)
    ref16_model = _copy.deepcopy(ref32.model).half()
    model_ph = _copy.deepcopy(ref16_model)
    yolo_head.patch_head(yolo_head.find_head(model_ph))
    p1h = []
    for f in (files[0], files[-1]):
        for size in sizes:
            x, _, _ = quiet(ref32.preprocess_image, f, size)
            with torch.no_grad():
                pr = ref16_model(x.half())[0]
                pm = model_ph(x.half())[0]
            same = pr.dtype == pm.dtype == torch.float32 and tuple(pr.shape) == tuple(pm.shape) and torch.equal(pr, pm)
            if not same and tuple(pr.shape) == tuple(pm.shape):
                dcol = (pr.float() != pm.float()).reshape(-1, pr.shape[-1]).sum(0).tolist()
                p1h.append({"file": os.path.basename(f), "size": size, "dtype_ref": str(pr.dtype), "dtype_mine": str(pm.dtype),
                            "differing_per_column": dcol, "max_abs": float((pr.float() - pm.float()).abs().max())})
            check(same, ("P1h half-model decode bit-identical", os.path.basename(f), size, p1h[-1] if p1h and not same else None))
            r_list = quiet(ref_general.non_max_suppression, pr.clone(), conf, 0.5)
            m_list = G.non_max_suppression(pm.clone(), conf, 0.5)
            check(len(r_list) == len(m_list) and all(tuple(r.shape) == tuple(m.shape) and torch.equal(r, m) for r, m in zip(r_list, m_list)),
                  ("P2h nms after the half-model decode", os.path.basename(f), size))
    REPORT["P1h_half_model_decode_diffs"] = p1h
    th_rep = []
    for f in (files[0], files[-1]):  # the same augmentation with the half model (half images through scale_img, float32 pred)
        x, _, _ = quiet(ref32.preprocess_image, f, sizes[0])
        with torch.no_grad():
            tr = ref16_model(x.half(), augment=True)[0]
            tm = TU.forward_augment(model_ph.forward_once, x.half(), gs=int(ref32.model.stride.max()))[0]
        same = tr.dtype == tm.dtype and tuple(tr.shape) == tuple(tm.shape) and torch.equal(tr, tm)
        th_rep.append({"file": os.path.basename(f), "rows": int(tr.shape[1]), "bit_identical": bool(same), "dtype_ref": str(tr.dtype),
                       "max_abs": float((tr.float() - tm.float()).abs().max()) if tuple(tr.shape) == tuple(tm.shape) else None})
        check(same, ("T forward_augment half model", th_rep[-1]))
    REPORT["T_forward_augment_half"] = th_rep
    del ref16_model, model_ph
    # P2c: multi-class predictions without keypoints (kpt_label=0, no = 5 + nc): class offset, `classes`, `multi_label`,
    # `agnostic` (utils/general.py:533-582) -- synthetic clustered boxes with 4 classes, both implementations on the same tensor
    g = torch.Generator(device="cpu").manual_seed(3)
    n_rows, ncls = 6000, 4
    ctr = torch.rand(40, 2, generator=g) * 600 + 20
    pick = torch.randint(0, 40, (2, n_rows), generator=g)
    xy = ctr[pick] + torch.randn(2, n_rows, 2, generator=g) * 6
    wh = torch.rand(2, n_rows, 2, generator=g) * 60 + 10
    mc = torch.cat((xy, wh, torch.rand(2, n_rows, 1, generator=g), torch.rand(2, n_rows, ncls, generator=g)), 2).to(ref32.device)
    for kw in (dict(), dict(multi_label=True), dict(agnostic=True), dict(classes=[1, 3]), dict(multi_label=True, classes=[0], agnostic=True)):
        r_mc = quiet(ref_general.non_max_suppression, mc.clone(), 0.3, 0.45, kpt_label=0, **kw)
        m_mc = G.non_max_suppression(mc.clone(), 0.3, 0.45, kpt_label=0, **kw)
        check(len(r_mc) == len(m_mc) and all(tuple(r.shape) == tuple(m.shape) and torch.equal(r, m) for r, m in zip(r_mc, m_mc))
              and sum(len(r) for r in r_mc) > 50, ("P2c multi-class nms", kw, [tuple(r.shape) for r in r_mc], [tuple(m.shape) for m in m_mc]))
    REPORT["P0_network_input_diffs"] = p0
    del mine_p
    REPORT["P1_decode"] = {"aten_cuda_mode": "bit-identical to the reference's CUDA fp32 decode",
                           "exact_mode_vs_aten_cuda": {"max_ulp": worst_ulp, "max_abs_diff_px_for_values_below_256": worst_px,
                                                       "ulp_histogram_0..7+": [int(v) for v in hist],
                                                       "identical_fraction": round(float(hist[0] / hist.sum()), 5)}}
    del model_p

    # ---- A: class swap, fp32 (near-exact, see the module docstring)
    mine32 = Mine(ckpt, "", sizes, conf, 0.5, half=False)
    ref_out, a_rep = {}, []
    for f in files:
        r, shape_r = quiet(ref32.detect_multi_scale, f)
        m, shape_m = mine32.detect_multi_scale(f)
        r, m = np.asarray(r, np.float64).reshape(-1, 7), np.asarray(m, np.float64).reshape(-1, 7)
        ref_out[f] = r
        rs = set(map(tuple, np.round(r[:, [0, 1, 2, 3, 5, 6]]).astype(np.int64).tolist()))
        ms_ = set(map(tuple, np.round(m[:, [0, 1, 2, 3, 5, 6]]).astype(np.int64).tolist()))
        exact = len(rs & ms_) / max(1, len(rs))
        near = min(match_boxes(r, m, 0.95), match_boxes(m, r, 0.95))
        a_rep.append({"file": os.path.basename(f), "reference": len(r), "libfdms": len(m), "rows_reproduced_exactly": round(exact, 4),
                      "matched_iou>=0.95": round(near, 4), "bit_identical": bool(rows_equal(m, r, f))})
        check(tuple(shape_r) == tuple(shape_m) and bool(rows_equal(m, r, f)), ("A class swap fp32", a_rep[-1]))
    REPORT["A_class_swap_fp32"] = a_rep

    if args.quick:
        print(json.dumps({"checks": N[0], "failures": [repr(x) for x in FAIL[:20]], "report": REPORT}))
        return 1 if FAIL else 0

    # ---- B: monkey patch of the reference's own class, fp32
    patched = ref_detector(False)
    patched.conf_thres = conf
    n_bound = fi.monkey_patch()
    fi.patch_model(patched.model)
    b_rep = []
    try:
        for f in files:
            r, _ = quiet(patched.detect_multi_scale, f)
            r = np.asarray(r, np.float64).reshape(-1, 7)
            rs = set(map(tuple, np.round(ref_out[f][:, [0, 1, 2, 3, 5, 6]]).astype(np.int64).tolist()))
            ms_ = set(map(tuple, np.round(r[:, [0, 1, 2, 3, 5, 6]]).astype(np.int64).tolist()))
            b_rep.append({"file": os.path.basename(f), "unpatched": len(ref_out[f]), "patched": len(r),
                          "rows_reproduced_exactly": round(len(rs & ms_) / max(1, len(rs)), 4), "bit_identical": bool(rows_equal(r, ref_out[f], f))})
            check(bool(rows_equal(r, ref_out[f], f)), ("B monkey patch fp32", b_rep[-1]))
    finally:
        fi.unpatch()
    REPORT["B_monkey_patch_fp32"], REPORT["B_rebound_names"] = b_rep, n_bound

    # ---- C: class swap against the reference's default CUDA mode (a half model; its pred is float32 with half-rounded
    # columns, everything after the head is float32 arithmetic)
    ref16 = ref_detector(True)
    ref16.conf_thres = conf
    mine16 = Mine(ckpt, "", sizes, conf, 0.5, half=True)
    agree = []
    for f in files:
        r, _ = quiet(ref16.detect_multi_scale, f)
        m, _ = mine16.detect_multi_scale(f)
        r, m = np.asarray(r, np.float64).reshape(-1, 7), np.asarray(m, np.float64).reshape(-1, 7)
        agree.append({"file": os.path.basename(f), "ref_half": len(r), "libfdms": len(m), "bit_identical": bool(rows_equal(m, r, f)),
                      "ref_boxes_matched_iou>=0.7": round(match_boxes(r, m, 0.7), 4), "libfdms_boxes_matched": round(match_boxes(m, r, 0.7), 4)})
        check(bool(rows_equal(m, r, f)), ("C class swap, half model", agree[-1]))
    REPORT["C_half_mode"] = agree

    # ---- C2: the production preprocessing (`use_api_preprocess=True`, what yolov7_face_multi_scale_dataframe_predict.py:109
    # constructs: PIL decode, zero-padded square, letterbox(auto=False); ppyp:345-378), class swap in both precisions
    api_rep = []
    for half in (False, True):
        ref_api = ref_detector(half, api=True)
        ref_api.conf_thres = conf
        mine_api = Mine(ckpt, "", sizes, conf, 0.5, use_api_preprocess=True, half=half)
        for f in files:
            r, shape_r = quiet(ref_api.detect_multi_scale, f)
            m, shape_m = mine_api.detect_multi_scale(f)
            r, m = np.asarray(r, np.float64).reshape(-1, 7), np.asarray(m, np.float64).reshape(-1, 7)
            api_rep.append({"file": os.path.basename(f), "half": half, "reference": len(r), "libfdms": len(m),
                            "bit_identical": bool(rows_equal(m, r, f))})
            check(tuple(shape_r) == tuple(shape_m) and bool(rows_equal(m, r, f)), ("C2 class swap, api preprocessing", api_rep[-1]))
        del ref_api, mine_api
    REPORT["C2_api_preprocess"] = api_rep

    # ---- D: unmodified detect.py, half precision, label files
    if not args.no_detect_py:
        import detect as ref_detect

        def run_detect(name):
            opt = argparse.Namespace(weights=ckpt, source=os.path.dirname(files[0]), img_size=sizes[0], conf_thres=conf, iou_thres=0.45,
                                     device="", view_img=False, save_txt=True, save_txt_tidl=False, save_bin=False, save_conf=True,
                                     save_crop=False, nosave=True, classes=None, agnostic_nms=False, augment=False, update=False,
                                     project=os.path.join(tmp, "runs"), name=name, exist_ok=True, line_thickness=3, hide_labels=False,
                                     hide_conf=False, kpt_label=5)
            with torch.no_grad():
                quiet(ref_detect.detect, opt)
            out = {}
            for p in sorted(glob.glob(os.path.join(tmp, "runs", name, "labels", "*.txt"))):
                out[os.path.basename(p)] = np.loadtxt(p, ndmin=2)
            return out

        base = run_detect("unpatched")
        fi.monkey_patch()
        orig_load = ref_detect.attempt_load

        def load_and_patch(*a, **k):
            m = orig_load(*a, **k)
            fi.patch_model(m)
            return m
        ref_detect.attempt_load = load_and_patch
        try:
            new = run_detect("patched")
        finally:
            ref_detect.attempt_load = orig_load
            fi.unpatch()
        rep = []
        for k in sorted(set(base) | set(new)):
            a, b = base.get(k, np.zeros((0, 6))), new.get(k, np.zeros((0, 6)))
            # rows: cls cx cy w h conf (normalised xywh): back to corner boxes for the IoU match
            ca = np.stack([a[:, 1] - a[:, 3] / 2, a[:, 2] - a[:, 4] / 2, a[:, 1] + a[:, 3] / 2, a[:, 2] + a[:, 4] / 2], 1) if len(a) else a[:, :4]
            cb = np.stack([b[:, 1] - b[:, 3] / 2, b[:, 2] - b[:, 4] / 2, b[:, 1] + b[:, 3] / 2, b[:, 2] + b[:, 4] / 2], 1) if len(b) else b[:, :4]
            rep.append({"file": k, "unpatched": len(a), "patched": len(b), "identical": bool(a.shape == b.shape and np.array_equal(a, b)),
                        "matched_iou>=0.7": round(match_boxes(ca, cb, 0.7), 4)})
        REPORT["D_detect_py"] = rep
        check(len(rep) > 0 and all(r["identical"] for r in rep), ("D detect.py label files", rep))

    # ---- E: timings on this GPU (wall clock around synchronised regions, best of 3)
    def best(fn, n=3):
        ts = []
        for _ in range(n):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return round(1e3 * min(ts), 3)
    f = syn
    quiet(ref16.detect_multi_scale, f); mine16.detect_multi_scale(f)
    REPORT["E_ms_per_image_1080p"] = {
        "reference_cuda_half": best(lambda: quiet(ref16.detect_multi_scale, f)),
        "libfdms_half": best(lambda: mine16.detect_multi_scale(f)),
        "reference_cuda_fp32": best(lambda: quiet(ref32.detect_multi_scale, f)),
        "libfdms_fp32": best(lambda: mine32.detect_multi_scale(f)),
    }
    bb = {}
    with torch.no_grad():
        for s in sizes:
            x, _, _ = quiet(ref16.preprocess_image, f, s)
            ref16.model(x)
            bb[str(s)] = best(lambda: ref16.model(x))
    REPORT["E_backbone_ms_per_image_half"] = bb

    print(json.dumps({"checks": N[0], "failures": [repr(x) for x in FAIL[:20]], "report": REPORT}))
    return 1 if FAIL else 0


if __name__ == "__main__":
    sys.exit(main())
