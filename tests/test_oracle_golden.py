"""Pins the CPU oracle (oracle/) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and against the third-party ops the reference calls.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from workloads import synth
from oracle import oracle as O
from tests import helpers as H


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- letterbox / pack (utils/datasets.py:873)
@pytest.mark.parametrize("i", range(len(H.LETTERBOX_SMALL)))
def test_letterbox_small(golden, i):
    g = golden("letterbox")
    h, w, ns, auto, scaleup, stride, seed = H.LETTERBOX_SMALL[i]
    img = synth.synth_image(h, w, seed)
    out, ratio, pad = O.letterbox(img, ns, auto=auto, scaleup=scaleup, stride=stride)
    ref = g["small%d_out" % i]
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)  # bit-exact vs cv2.resize + copyMakeBorder
    meta = g["small_meta"][i]
    assert ratio[0] == meta[8] and pad == (meta[9], meta[10])
    f16 = O.pack_chw(out, flip=True, dtype=np.float16)
    f32 = O.pack_chw(out, flip=True, dtype=np.float32)
    assert sha(f16.view(np.uint16)) == str(g["small%d_f16_sha" % i])
    assert sha(f32) == str(g["small%d_f32_sha" % i])
    if i in (0, 3):
        assert np.array_equal(f16.view(np.uint16), g["small%d_f16" % i])
        assert np.array_equal(f32, g["small%d_f32" % i])


def test_letterbox_per_channel_colour(golden):
    g = golden("letterbox_color")
    for i, (h, w, ns0, ns1, auto, c0, c1, c2, seed) in enumerate(g["meta"]):
        img = synth.synth_image(int(h), int(w), int(seed))
        ns = int(ns0) if ns0 == ns1 else (int(ns0), int(ns1))
        out, ratio, pad = O.letterbox(img, ns, color=(int(c0), int(c1), int(c2)), auto=bool(auto))
        assert np.array_equal(out, g["c%d_out" % i])
        assert (ratio[0], ratio[1], pad[0], pad[1]) == tuple(g["c%d_ratio_pad" % i])


@pytest.mark.parametrize("i", range(len(H.LETTERBOX_FULL)))
def test_letterbox_full_digest(golden, i):
    g = golden("letterbox")
    h, w, ns, auto, stride, seed = H.LETTERBOX_FULL[i]
    img = synth.synth_image(h, w, seed)
    out, _, _ = O.letterbox(img, ns, auto=auto, stride=stride)
    assert out.shape[:2] == tuple(g["full_meta"][i][6:8])
    assert sha(out) == str(g["full_sha_u8"][i])
    assert sha(O.pack_chw(out, True, np.float16)) == str(g["full_sha_f16"][i])


@pytest.mark.parametrize("i", range(len(H.API_CASES)))
def test_api_preprocess(golden, i):
    g = golden("letterbox")
    h, w, size, seed = H.API_CASES[i]
    rgb = synth.synth_image(h, w, seed)
    chw = O.preprocess_api(rgb, size, 32)
    assert sha(chw) == str(g["api%d_sha" % i])
    if "api%d_out" % i in g:
        assert np.array_equal(chw, g["api%d_out" % i])


def test_resize_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for (sh, sw), (dh, dw) in [((211, 173), (64, 53)), ((53, 37), (143, 100)), ((480, 640), (960, 1280)),
                               ((1, 9), (4, 31)), ((300, 300), (300, 300)), ((719, 1279), (360, 640))]:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(O.resize_linear_u8(img, dh, dw), ref)


# ---------------------------------------------------------------- head decode (models/yolo.py:267-308)
def _ulps(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia); ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return np.abs(ia - ib)


def test_decode_tiny_model(golden):
    g = golden("head")
    levels = [g["tiny_raw%d" % i] for i in range(3)]
    pred = O.decode_head(levels, g["tiny_strides"], g["tiny_anchors_px"])
    ref = g["tiny_pred"]
    assert pred.shape == ref.shape
    # 1e-4 px (north_star); libm expf vs ATen's vectorised exp differ by <= 1 ulp in the sigmoid
    assert np.max(np.abs(pred - ref)) <= 1e-4
    # (ulp distance is not meaningful where 2*sigma-0.5+grid cancels; the bound is absolute)
    assert np.mean(_ulps(pred, ref) == 0) > 0.9


@pytest.mark.parametrize("name", list(H.CROWD_HEADS))
def test_decode_and_nms_crowd(golden, name):
    g = golden("head")
    in_h, in_w, n_faces, seed = H.CROWD_HEADS[name]
    levels = H.crowd_levels(in_h, in_w, n_faces, seed)
    pred = O.decode_head(levels, synth.P5_STRIDES, synth.P5_ANCHORS)
    ref_head = g[name + "_pred_b0_head"]
    assert np.max(np.abs(pred[0, :64] - ref_head)) <= 1e-4
    for conf, iou in ((0.5, 0.5), (0.6, 0.3), (0.25, 0.45)):
        res = O.non_max_suppression(pred, conf, iou, O.TV_CPU)
        for b in range(2):
            ref = g["%s_nms_c%d_i%d_b%d" % (name, int(conf * 100), int(iou * 100), b)]
            assert res[b].shape == ref.shape, (name, conf, iou, b)
            assert np.max(np.abs(res[b] - ref)) <= 1e-4 if ref.size else True


HEAD_VARIANTS = ("detect_nc3", "detect_nc80", "idetect_nc1", "idetect_pose", "detect_face", "ikeypoint_k3")


@pytest.mark.parametrize("name", HEAD_VARIANTS)
def test_decode_head_variants(golden, name):
    """Detect / IDetect / IKeypoint of models/yolo.py with other na / nc / nkpt than the face head: the oracle's general
    decode against the reference modules' own eval-mode output (tests/golden/make_golden.py::gen_heads)."""
    g = golden("heads")
    nc, nkpt, na, _ = (int(v) for v in g[name + "__meta"])
    no = 5 + nc + 3 * nkpt
    raw = [g["%s__raw%d" % (name, i)] for i in range(3)]
    pred = O.decode_head_general(raw, (8, 16, 32), g[name + "__anchors_px"], na, no, nkpt)
    ref = g[name + "__pred"]
    assert pred.shape == ref.shape
    # 1e-4 px (north_star), or 4 ulp where a box is so large (anchors up to 200 px x 4) that 1e-4 px is less than that:
    # ATen's vectorised CPU exp (Sleef) and the correctly rounded one differ by <= 1 ulp in the sigmoid, (2s)^2 doubles
    # the relative error and the product with the anchor rounds once more (seen: 1.2e-4 at 379 px)
    assert np.all(np.abs(pred - ref) <= np.maximum(1e-4, 4 * np.spacing(np.abs(ref))))
    assert np.mean(pred == ref) > 0.9
    if nkpt == 5 and na == 3:  # the face layout: the specialised restatement agrees bit for bit
        assert np.array_equal(O.decode_head(raw, (8, 16, 32), g[name + "__anchors_px"]), pred)


@pytest.mark.parametrize("name", ("detect_face", "idetect_pose", "ikeypoint_k3", "idetect_nc1"))
def test_decode_head_half_model(golden, name):
    """A HALF model (the reference's default on a GPU): float32 pred with half-rounded columns, bit for bit against the
    reference modules run on half tensors (tests/golden/make_golden.py::gen_heads_half)."""
    g = golden("heads_half")
    nc, nkpt, na, _ = (int(v) for v in g[name + "__meta"])
    raw = [g["%s__raw%d" % (name, i)] for i in range(3)]
    assert raw[0].dtype == np.float16 and g[name + "__pred"].dtype == np.float32
    pred = O.decode_head_half(raw, (8, 16, 32), g[name + "__anchors_px"], na, 5 + nc + 3 * nkpt, nkpt)
    assert np.array_equal(pred, g[name + "__pred"])


def test_nms_on_reference_pred(golden):
    """filter + NMS from the reference's own decoded pred: bit-exact rows."""
    g = golden("head")
    pred = g["tiny_pred"]
    thr = float(g["tiny_conf_thres"])
    for iou in (0.45, 0.3):
        res = O.non_max_suppression(pred, thr, iou, O.TV_CPU)
        for b in range(2):
            ref = g["tiny_nms_iou%d_b%d" % (int(iou * 100), b)]
            assert ref.shape[0] > 10
            assert np.array_equal(res[b], ref)


# ---------------------------------------------------------------- torchvision.ops.nms known answers
@pytest.mark.parametrize("name", ["rand", "int", "exact", "crowd", "one"])
def test_nms_golden(golden, name):
    g = golden("nms")
    boxes, scores = g[name + "_boxes"], g[name + "_scores"]
    for t in g[name + "_thrs"]:
        ref = g["%s_keep_%d" % (name, int(round(t * 1e6)))]
        got = O.nms(boxes, scores, float(t), O.TV_CPU)
        assert np.array_equal(got, ref), (name, t)


@pytest.mark.parametrize("case", range(5))
def test_nms_ncnn_mode_golden(golden, case):
    """Oracle mode 2 against the reference's own ncnn-demo C++ NMS (fixture: make_golden.py ncnn)."""
    g = golden("ncnn")
    got = O.nms(g["b%d" % case], g["s%d" % case], float(g["t%d" % case]), O.NCNN)
    assert np.array_equal(got, g["k%d" % case])


def test_nms_matches_torchvision_live():
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 257, 2000):
        xy = rng.integers(0, 80, (n, 2)).astype(np.float32); wh = rng.integers(1, 30, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = np.round(rng.uniform(0, 1, n), 2).astype(np.float32)
        for t in (0.3, 0.5, 1 / 3):
            ref = tv.ops.nms(torch.from_numpy(boxes).reshape(-1, 4), torch.from_numpy(scores), t).numpy()
            assert np.array_equal(O.nms(boxes, scores, t, O.TV_CPU), ref)


def test_nms_modes_differ_only_near_threshold():
    """TV_CUDA (fused area) vs TV_CPU formulas agree except for IoUs within an ulp of the threshold."""
    rng = np.random.default_rng(12)
    xy = rng.uniform(0, 300, (1500, 2)); wh = rng.uniform(5, 80, (1500, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    scores = rng.uniform(0, 1, 1500).astype(np.float32)
    assert np.array_equal(O.nms(boxes, scores, 0.45, O.TV_CPU), O.nms(boxes, scores, 0.45, O.TV_CUDA))
    # exact-threshold: (double)rn32(1/3) > 1/3 suppresses on the CPU kernel, the fp32 compare does not
    b = np.array([[0, 0, 2, 1], [1, 0, 3, 1]], np.float32); s = np.array([0.9, 0.8], np.float32)
    assert list(O.nms(b, s, 1 / 3, O.TV_CPU)) == [0]
    assert list(O.nms(b, s, 1 / 3, O.TV_CUDA)) == [0, 1]


# ---------------------------------------------------------------- rescale (utils/general.py:374, ppyp:122)
def test_rescale_golden(golden):
    g = golden("rescale")
    for i, (ih, iw, h0, w0) in enumerate(g["meta"]):
        b = g["c%d_boxes" % i]; k = g["c%d_kpts" % i]
        assert np.array_equal(O.scale_coords((ih, iw), b.copy(), (h0, w0)), g["c%d_std_boxes" % i])
        assert np.array_equal(O.scale_coords((ih, iw), k.copy(), (h0, w0), kpt_label=5, step=3), g["c%d_std_kpts" % i])
        assert np.array_equal(O.scale_coords_api((ih, iw), b.copy(), (h0, w0)), g["c%d_api_boxes" % i])
        assert np.array_equal(O.scale_coords_api((ih, iw), b.copy(), (h0, w0), do_round=True), g["c%d_api_round" % i])


def test_rescale_ratio_pad_golden(golden):
    """scale_coords(ratio_pad=...) (utils/general.py:379-383): tuple gain, test.py's (h/h0, w/w0) and a scalar gain."""
    g = golden("rescale_rp")
    for i, (ih, iw, h0, w0) in enumerate(g["meta"]):
        for j in range(3):
            gain, px, py = g["c%d_%d_rp" % (i, j)]
            rp = ((gain, gain), (px, py)) if j < 2 else (gain, (px, py))
            b = g["c%d_%d_boxes" % (i, j)]; k = g["c%d_%d_kpts" % (i, j)]
            assert np.array_equal(O.scale_coords((ih, iw), b.copy(), (h0, w0), rp), g["c%d_%d_std_boxes" % (i, j)])
            assert np.array_equal(O.scale_coords((ih, iw), k.copy(), (h0, w0), rp, kpt_label=5, step=3),
                                  g["c%d_%d_std_kpts" % (i, j)])

# ---------------------------------------------------------------- merge (multi_scale_face_detector.py:168-240)
def test_merge_golden(golden):
    g = golden("merge")
    for ci in range(5):
        dets = g["m%d_dets" % ci]
        n_scales, iou = g["m%d_meta" % ci]
        assert np.array_equal(O.scale_weights(dets, int(n_scales)), g["m%d_weights" % ci])
        assert np.array_equal(O.weighted_nms(dets, int(n_scales), float(iou)), g["m%d_out" % ci])


# ---------------------------------------------------------------- end to end (msfd:109-166,242-288)
@pytest.mark.parametrize("case", H.E2E_CASES, ids=[c[0] for c in H.E2E_CASES])
def test_e2e_golden(golden, case):
    g = golden("e2e")
    name, h, w, sizes, n_faces, conf, iou, api, seed = case
    bgr = synth.synth_image(h, w, seed)
    heads, shapes = H.e2e_heads(h, w, sizes, n_faces, api, seed)
    assert np.array_equal(np.array(shapes), g[name + "_shapes"])
    # preprocessing: reference model inputs (fp32 on CPU) by digest
    for si, size in enumerate(sizes):
        if api:
            chw = O.preprocess_api(bgr[:, :, ::-1], size, 32)  # PIL gives RGB; no channel flip (ppyp:375)
            x = chw.astype(np.float32) / np.float32(255.0)
        else:
            lb, _, _ = O.letterbox(bgr, size, auto=True, stride=32)
            x = O.pack_chw(lb, flip=True, dtype=np.float32)
        assert x.shape[1:] == tuple(shapes[si])
        assert sha(x[None]) == str(g["%s_input%d_sha" % (name, si)])
    final, per_scale = O.detect_from_heads([[l[None] for l in heads[s]] for s in shapes], shapes, (h, w, 3),
                                           synth.P5_STRIDES, synth.P5_ANCHORS, conf, iou, O.TV_CPU)
    for si in range(len(sizes)):
        ref = g["%s_scale%d" % (name, si)]
        assert per_scale[si].shape == ref.shape
        assert np.array_equal(per_scale[si], ref)
    ref = g[name + "_final"]
    assert final.shape == ref.shape
    assert np.array_equal(final, ref)
    if n_faces:
        assert ref.shape[0] >= 3


def test_nms_other_keypoint_counts(golden):
    """non_max_suppression on 6- and 12-column predictions (kpt_label 0 / 2, nc = 1): reference golden."""
    g, h = golden("nms6"), golden("head")
    thr = float(g["thr"])
    for cols in (6, 12):
        res = O.non_max_suppression(np.ascontiguousarray(h["tiny_pred"][..., :cols]), thr, 0.45, O.TV_CPU)
        for b in range(2):
            assert np.array_equal(res[b], g["c%d_b%d" % (cols, b)])


# ---- "next" row: test-time augmentation around the backbone (models/yolo.py:363-417, utils/torch_utils.py:247-257)
TTA_TOL = 2.4e-7  # 2 ulp at 1.0: ATen's build contracts the 4-tap sum differently; inputs are in [0, 1]


def test_tta_scale_img_golden(golden):
    g = golden("tta")
    for i in range(5):
        ratio, flip, gs = g["s%d_args" % i]
        got = O.scale_img(g["x"], float(ratio), gs=int(gs), flip=int(flip) or None)
        assert got.shape == g["s%d" % i].shape
        assert np.max(np.abs(got - g["s%d" % i])) <= TTA_TOL
    got = O.scale_img(g["x"], 0.67, same_shape=True)
    assert got.shape == g["same_shape"].shape and np.max(np.abs(got - g["same_shape"])) <= TTA_TOL


def test_tta_descale_pred_golden(golden):
    g = golden("tta")
    ys = []
    for i, (si, fi) in enumerate(zip([1, 0.83, 0.67], [None, 3, None])):
        ys.append(O.descale_pred(g["p"], fi, si, (48, 80)))
        assert np.array_equal(ys[-1], g["d%d" % i])
    assert np.array_equal(O.descale_pred(g["p"], 2, 0.83, (48, 80)), g["d_ud"])
    assert np.array_equal(np.concatenate(ys, 1), g["cat"])


def test_nms_with_apriori_labels_golden(golden):
    """utils/general.py:533-541 (autolabelling) on the 6-column prediction; fixture: make_golden.py labels."""
    g, h = golden("labels"), golden("head")
    pred = np.ascontiguousarray(h["tiny_pred"][..., :6])
    res = O.non_max_suppression(pred, float(g["thr"]), 0.45, O.TV_CPU, labels=[g["l0"], g["l1"]])
    for b in range(2):
        assert np.array_equal(res[b], g["r%d" % b])


# ---------------------------------------------------------------- the oracle against the live reference
@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_matches_live_reference_on_random_inputs(seed):
    """Only where the reference tree exists (the build container): tests/ref_live_check.py imports the UNMODIFIED
    reference in a subprocess and compares ~200 seeded random calls of its functions (letterbox, api preprocessing, every head class incl. half models,
    IKeypoint decode, non_max_suppression, scale_coords*, weighted_nms, scale_img, _descale_pred) with the oracle."""
    import json
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/utils"):
        pytest.skip("reference tree not present (GPU box / deployment): the committed fixtures pin the oracle there")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "ref_live_check.py"), str(seed)], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, (r.stdout[-500:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    assert r.returncode == 0 and not res["failures"] and res["checks"] > 150, res
