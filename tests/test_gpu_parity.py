"""Parity of the CUDA path (libfdms through its C ABI) against the CPU oracle and the reference golden
fixtures.  GPU only (`-m gpu`).  Tolerances (BASELINE.json north_star):
  pyramid pixels        bit-exact here (north_star allows 1 LSB vs cv2 INTER_LINEAR)
  candidate counts/sets bit-exact
  decoded values        <= 1e-4 px absolute
  NMS keep lists        bit-exact
"""
import hashlib

import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import _lib, ops
from workloads import synth
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_PX = 1e-4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# =============================================================================== pyramid
def _pyr(img, geoms, dtype, layout="chw", flip=True, virt=None):
    src = gpu(img[None])
    # shared memory the kernel reads without having written it must not go unnoticed: start from a poisoned SM
    ops.debug_poison_smem(0xFF if (img.shape[0] + img.shape[1]) & 1 else 0x7F)
    outs = ops.pyramid(src, geoms, out_dtype=dtype, layout=layout, flip_rb=flip, virt_hw=virt)
    torch.cuda.synchronize()
    return [o[0].cpu().numpy() for o in outs]


@pytest.mark.parametrize("i", range(len(H.LETTERBOX_SMALL)))
def test_letterbox_small_u8_and_pack(golden, i):
    g = golden("letterbox")
    h, w, ns, auto, scaleup, stride, seed = H.LETTERBOX_SMALL[i]
    img = synth.synth_image(h, w, seed)
    geom = _lib.letterbox_geometry(h, w, ns, auto=auto, scaleup=scaleup, stride=stride)
    ref = g["small%d_out" % i]
    assert (geom.dst_h, geom.dst_w) == ref.shape[:2]
    (u8,) = _pyr(img, [geom], torch.uint8, layout="hwc", flip=False)
    assert np.array_equal(u8, ref)  # == cv2.resize + copyMakeBorder of the reference letterbox
    (f16,) = _pyr(img, [geom], torch.float16)
    (f32,) = _pyr(img, [geom], torch.float32)
    assert sha(f16.view(np.uint16)) == str(g["small%d_f16_sha" % i])
    assert sha(f32) == str(g["small%d_f32_sha" % i])


@pytest.mark.parametrize("i", range(len(H.LETTERBOX_FULL)))
def test_letterbox_full_digest(golden, i):
    g = golden("letterbox")
    h, w, ns, auto, stride, seed = H.LETTERBOX_FULL[i]
    img = synth.synth_image(h, w, seed)
    geom = _lib.letterbox_geometry(h, w, ns, auto=auto, stride=stride)
    (u8,) = _pyr(img, [geom], torch.uint8, layout="hwc", flip=False)
    assert u8.shape[:2] == tuple(g["full_meta"][i][6:8])
    assert sha(u8) == str(g["full_sha_u8"][i])
    (f16,) = _pyr(img, [geom], torch.float16)
    assert sha(f16) == str(g["full_sha_f16"][i])


def test_pyramid_three_scales_batch_one_launch(golden):
    """cfg2 shape: all scales x a batch in ONE launch equals the per-scale reference letterbox."""
    g = golden("letterbox")
    h, w, seed = 1080, 1920, 100
    imgs = np.stack([synth.synth_image(h, w, seed), synth.synth_image(h, w, seed + 1)])
    geoms = [_lib.letterbox_geometry(h, w, s, auto=True, stride=32) for s in (640, 960, 1280)]
    outs = ops.pyramid(gpu(imgs), geoms, out_dtype=torch.float16)
    torch.cuda.synchronize()
    assert [tuple(o.shape) for o in outs] == [(2, 3, 384, 640), (2, 3, 544, 960), (2, 3, 736, 1280)]
    for s in range(3):
        assert sha(outs[s][0].cpu().numpy()) == str(g["full_sha_f16"][s])  # LETTERBOX_FULL[0..2] share seed 100
        lb, _, _ = O.letterbox(imgs[1], (640, 960, 1280)[s], auto=True, stride=32)
        assert np.array_equal(outs[s][1].cpu().numpy().view(np.uint16), O.pack_chw(lb, True, np.float16).view(np.uint16))


@pytest.mark.parametrize("i", range(len(H.API_CASES)))
def test_api_preprocess(golden, i):
    g = golden("letterbox")
    h, w, size, seed = H.API_CASES[i]
    rgb = synth.synth_image(h, w, seed)
    s = max(h, w)
    geom = _lib.letterbox_geometry(s, s, size, auto=False, stride=32)
    (chw,) = _pyr(rgb, [geom], torch.uint8, layout="chw", flip=False, virt=(s, s))
    assert sha(chw) == str(g["api%d_sha" % i])


@pytest.mark.parametrize("hw,sizes", [((384, 216), (128, 192, 256)), ((1920, 1080), (640, 960, 1280)), ((216, 384), (128, 192, 256)),
                                      ((360, 640), (320, 160, 480)), ((720, 1280), (640, 960, 1280)), ((96, 160), (160,))])
def test_letterbox_exact_factors_with_column_padding(hw, sizes):
    """Portrait (column-padded) and landscape letterboxes at exact 3:1 / 2:1 / 3:2 factors: the direct tiles incl.
    their border column pairs, fp16 CHW with the channel flip like the hot path, bit-exact against the oracle."""
    h, w = hw
    img = synth.synth_image(h, w, 11 + w)
    geoms = [_lib.letterbox_geometry(h, w, size, auto=True, stride=32) for size in sizes]
    outs = _pyr(img, geoms, torch.float16)
    for size, got in zip(sizes, outs):
        lb, _, _ = O.letterbox(img, size, auto=True, stride=32)
        assert np.array_equal(got.view(np.uint16), O.pack_chw(lb, True, np.float16).view(np.uint16)), size


@pytest.mark.parametrize("hw,sizes", [((120, 160), (80, 32, 96)), ((1080, 1920), (640, 960, 1280)), ((90, 160), (32, 80)),
                                      ((640, 360), (320, 160, 96)), ((1920, 1080), (640, 1280)), ((768, 1366), (640,))])
def test_api_mode_integer_scales_direct_tiles(hw, sizes):
    """api mode (virtual zero-padded square).  Landscape images whose square is an integer multiple of the model size:
    the 2:1 and odd:1 scales run as direct tiles with the virtual rows written as zeros, the others staged.  Portrait
    and odd-width images: virtual zero COLUMNS and rows that do not start on 16 bytes (shifted bulk staging).  All
    scales in one launch, fp16 CHW like the hot path, bit-exact against the oracle's preprocess_api + pack."""
    h, w = hw
    img = synth.synth_image(h, w, 7 + h)
    s = max(h, w)
    geoms = [_lib.letterbox_geometry(s, s, size, auto=False, stride=32) for size in sizes]
    outs = _pyr(img, geoms, torch.float16, layout="chw", flip=False, virt=(s, s))
    for size, got in zip(sizes, outs):
        ref = O.preprocess_api(img, size).astype(np.float32) / np.float32(255)
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint16), ref.astype(np.float16).view(np.uint16)), size


@pytest.mark.parametrize("shape", [((211, 173), (64, 53)), ((53, 37), (143, 100)), ((480, 640), (960, 1280)),
                                   ((1, 9), (4, 31)), ((300, 300), (300, 300)), ((719, 1279), (360, 640)),
                                   ((2160, 3840), (360, 640)), ((33, 2000), (17, 40))])
def test_resize_ragged_shapes_vs_oracle(shape):
    """Arbitrary (mis-aligned pitch, up- and down-scale, huge ratio) sources through the scalar staging path."""
    (sh, sw), (dh, dw) = shape
    img = np.random.default_rng(sh * 7 + dw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    geom = _lib.Geom()
    geom.dst_h, geom.dst_w, geom.unpad_h, geom.unpad_w = dh, dw, dh, dw
    geom.resize = 1
    (u8,) = _pyr(img, [geom], torch.uint8, layout="hwc", flip=False)
    assert np.array_equal(u8, O.resize_linear_u8(img, dh, dw))


def test_pyramid_strided_rows():
    """Source given as a view with a padded row pitch (16-byte aligned and not)."""
    for pad in (16, 5):
        full = np.random.default_rng(pad).integers(0, 256, (120, 200 + pad, 3), dtype=np.uint8)
        t = gpu(full)[None][:, :, :200, :]
        geom = _lib.letterbox_geometry(120, 200, 96, auto=True, stride=32)
        (out,) = ops.pyramid(t, [geom], out_dtype=torch.uint8, layout="hwc", flip_rb=False)
        ref, _, _ = O.letterbox(full[:, :200], 96, auto=True, stride=32)
        assert np.array_equal(out[0].cpu().numpy(), ref)


# =============================================================================== decode
def _crowd(name, dtype=np.float32):
    in_h, in_w, n_faces, seed = H.CROWD_HEADS[name]
    return [l.astype(dtype) for l in H.crowd_levels(in_h, in_w, n_faces, seed)]


@pytest.mark.parametrize("name", list(H.CROWD_HEADS))
def test_decode_dense_vs_oracle_and_reference(golden, name):
    g = golden("head")
    levels = _crowd(name)
    pred = ops.decode_dense([gpu(l) for l in levels], synth.P5_STRIDES, synth.P5_ANCHORS).cpu().numpy()
    ref = O.decode_head(levels, synth.P5_STRIDES, synth.P5_ANCHORS)
    assert pred.shape == ref.shape
    assert np.max(np.abs(pred - ref)) <= TOL_PX
    assert np.mean(pred == ref) > 0.999  # same correctly-rounded sigmoid: essentially bit identical
    assert np.max(np.abs(pred[0, :64] - g[name + "_pred_b0_head"])) <= TOL_PX  # reference IKeypoint.forward


def test_decode_dense_tiny_model_fp16_and_fp32(golden):
    g = golden("head")
    levels = [g["tiny_raw%d" % i] for i in range(3)]
    pred = ops.decode_dense([gpu(l) for l in levels], g["tiny_strides"], g["tiny_anchors_px"]).cpu().numpy()
    assert np.max(np.abs(pred - g["tiny_pred"])) <= TOL_PX
    lv16 = [l.astype(np.float16) for l in levels]
    p16 = ops.decode_dense([gpu(l) for l in lv16], g["tiny_strides"], g["tiny_anchors_px"]).cpu().numpy()
    ref16 = O.decode_head(lv16, g["tiny_strides"], g["tiny_anchors_px"])
    assert np.max(np.abs(p16 - ref16)) <= TOL_PX


def _sorted_cands(c, seg):
    n = int(c.count[seg].item())
    rows = c.row[seg, :n].cpu().numpy()
    cand = c.cand[seg, :n].cpu().numpy()
    o = np.argsort(rows, kind="stable")
    return cand[o], rows[o]


@pytest.mark.parametrize("name", list(H.CROWD_HEADS))
@pytest.mark.parametrize("conf", [0.5, 0.25, 0.6, 0.01])
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_decode_filter_counts_and_sets(name, conf, dtype):
    levels = _crowd(name, dtype)
    ref_pred = O.decode_head(levels, synth.P5_STRIDES, synth.P5_ANCHORS)
    c = ops.decode_filter([[gpu(l) for l in levels]], synth.P5_STRIDES, synth.P5_ANCHORS, conf, cap=ref_pred.shape[1])
    torch.cuda.synchronize()
    for b in range(ref_pred.shape[0]):
        rc, rr = O.filter_xyxy(ref_pred[b], conf)
        cand, rows = _sorted_cands(c, b)
        assert len(rows) == len(rr)          # candidate count: bit exact
        assert np.array_equal(rows, rr)      # candidate set: bit exact
        assert np.max(np.abs(cand - rc), initial=0) <= TOL_PX
        assert np.array_equal(cand[:, 4], rc[:, 4])  # conf drives the NMS order: must be identical


def test_decode_filter_multi_scale_segments():
    """Three scales x batch 2 in one launch: segment g = b * n_scales + s."""
    shapes = [(96, 160), (128, 224), (192, 320)]
    sl = []
    for si, (ih, iw) in enumerate(shapes):
        sl.append(H.crowd_levels(ih, iw, 10 + 5 * si, 40 + si))
    c = ops.decode_filter([[gpu(l) for l in lv] for lv in sl], synth.P5_STRIDES, synth.P5_ANCHORS, 0.4, cap=2048)
    torch.cuda.synchronize()
    for si, lv in enumerate(sl):
        pred = O.decode_head(lv, synth.P5_STRIDES, synth.P5_ANCHORS)
        for b in range(2):
            rc, rr = O.filter_xyxy(pred[b], 0.4)
            cand, rows = _sorted_cands(c, b * 3 + si)
            assert np.array_equal(rows, rr) and len(rr) > 0
            assert np.max(np.abs(cand - rc)) <= TOL_PX


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
def test_decode_filter_boxes_only_equals_full_decode(dtype):
    """FDMS_DECODE_BOXES_ONLY: same candidate sets, columns 0..5 bit-identical to the full decode, landmark columns
    untouched (the multi-scale path drops them: multi_scale_face_detector.py:151-153)."""
    sl = [[l.astype(dtype) for l in H.crowd_levels(ih, iw, 12, 50 + si)]
          for si, (ih, iw) in enumerate([(96, 160), (192, 320)])]
    dev_levels = [[gpu(l) for l in lv] for lv in sl]
    full = ops.decode_filter(dev_levels, synth.P5_STRIDES, synth.P5_ANCHORS, 0.3, cap=2048)
    lite = ops.Candidates(full.n_seg, full.cap, full.cand.device)
    lite.cand.fill_(-7.0)
    ops.decode_filter(dev_levels, synth.P5_STRIDES, synth.P5_ANCHORS, 0.3, cap=2048, out=lite, landmarks=False)
    torch.cuda.synchronize()
    assert np.array_equal(full.count.cpu().numpy(), lite.count.cpu().numpy())
    for seg in range(full.n_seg):
        cf, rf = _sorted_cands(full, seg)
        cl, rl = _sorted_cands(lite, seg)
        assert len(rf) > 0 and np.array_equal(rf, rl)
        assert np.array_equal(cf[:, :6], cl[:, :6])
        assert np.all(cl[:, 6:] == -7.0)


def test_decode_filter_overflow_is_reported():
    levels = _crowd("crowdA")
    c = ops.decode_filter([[gpu(l) for l in levels]], synth.P5_STRIDES, synth.P5_ANCHORS, 0.01, cap=16)
    cnt = c.count.cpu().numpy()
    pred = O.decode_head(levels, synth.P5_STRIDES, synth.P5_ANCHORS)
    assert cnt[0] == len(O.filter_xyxy(pred[0], 0.01)[1]) > 16  # total is still exact; caller sees count > cap


def test_filter_pred_matches_oracle(golden):
    g = golden("head")
    pred = g["tiny_pred"]
    thr = float(g["tiny_conf_thres"])
    c = ops.filter_pred(gpu(pred), thr, cap=pred.shape[1])
    for b in range(2):
        rc, rr = O.filter_xyxy(pred[b], thr)
        cand, rows = _sorted_cands(c, b)
        assert np.array_equal(rows, rr)
        assert np.array_equal(cand, rc)  # same fp32 inputs: bit exact


# =============================================================================== NMS
@pytest.mark.parametrize("name", ["rand", "int", "exact", "crowd", "one"])
def test_nms_golden_torchvision_cpu(golden, name):
    g = golden("nms")
    boxes, scores = gpu(g[name + "_boxes"]), gpu(g[name + "_scores"])
    for t in g[name + "_thrs"]:
        ref = g["%s_keep_%d" % (name, int(round(t * 1e6)))]
        got = ops.nms(boxes, scores, float(t), ops.IOU_TV_CPU).cpu().numpy()
        assert np.array_equal(got, ref), (name, t)


@pytest.mark.parametrize("mode", [ops.IOU_TV_CPU, ops.IOU_TV_CUDA])
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 257, 2000, 5000, 9000])
def test_nms_vs_oracle_sizes(mode, n):
    rng = np.random.default_rng(n)
    xy = rng.integers(0, 120, (n, 2)).astype(np.float32)
    wh = rng.integers(1, 40, (n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], 1)
    scores = np.round(rng.uniform(0, 1, n), 2).astype(np.float32)  # many exact ties
    for t in (0.3, 0.5, 1 / 3):
        ref = O.nms(boxes, scores, t, mode)
        got = ops.nms(gpu(boxes), gpu(scores), t, mode).cpu().numpy()
        assert np.array_equal(got, ref), (n, t)


@pytest.mark.parametrize("case", range(5))
def test_nms_ncnn_mode_golden(golden, case):
    """FDMS_IOU_NCNN against the reference's own C++ (cpp/yolov7-face-ncnn/src/yolov7face.cpp:35-74, fixture made by
    tests/golden/make_golden.py ncnn): +1 pixel extents, `>=`; 400 / 1500 / 3000 boxes take the fused, the sorted and
    the blocked-greedy / bit-matrix kernels."""
    g = golden("ncnn")
    b, s, t, ref = g["b%d" % case], g["s%d" % case], float(g["t%d" % case]), g["k%d" % case]
    got = ops.nms(gpu(b), gpu(s), t, ops.IOU_NCNN).cpu().numpy()
    assert np.array_equal(got, ref)
    # many segments at once (>= 8 segments: blocked greedy kernel for the large ones)
    n = len(s)
    rows = np.zeros((9, n, 6), np.float32)
    rows[:, :, :4], rows[:, :, 4] = b, s
    keep, n_keep = ops.nms_segments(gpu(rows), gpu(np.full((9,), n, np.int32)), n, t, ops.IOU_NCNN, max_nms=0)
    torch.cuda.synchronize()
    for seg in (0, 8):
        assert np.array_equal(keep[seg, :int(n_keep[seg])].cpu().numpy(), ref)


def test_nms_ncnn_mode_vs_oracle_with_ties():
    rng = np.random.default_rng(5)
    for n in (65, 700, 2500):
        xy = rng.integers(0, 150, (n, 2)).astype(np.float32)
        wh = rng.integers(0, 40, (n, 2)).astype(np.float32)  # zero-size boxes have area 1 here
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = np.round(rng.uniform(0, 1, n), 2).astype(np.float32)
        for t in (0.25, 0.5, 0.0):
            assert np.array_equal(ops.nms(gpu(boxes), gpu(scores), t, ops.IOU_NCNN).cpu().numpy(), O.nms(boxes, scores, t, O.NCNN))


def test_nms_matches_torchvision_cuda_kernel():
    """The TV_CUDA flavour against the real torchvision sm_100 kernel on the same box."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(77)
    for n in (300, 4000):
        xy = rng.uniform(0, 400, (n, 2)); wh = rng.uniform(4, 90, (n, 2))
        boxes = gpu(np.concatenate([xy, xy + wh], 1).astype(np.float32))
        scores = gpu(rng.uniform(0, 1, n).astype(np.float32))
        for t in (0.3, 0.45, 0.5):
            ref = tv.ops.nms(boxes, scores, t).cpu().numpy()
            got = ops.nms(boxes, scores, t, ops.IOU_TV_CUDA).cpu().numpy()
            assert np.array_equal(got, ref), (n, t)


def test_nms_ties_and_on_threshold_pairs_match_torchvision_cuda_kernel():
    """SURVEY B7's open point, against the real sm_100 kernel: quantised scores (exact ties -> the stable sort must keep
    the lower index first) and integer boxes whose IoU lies EXACTLY on the threshold (0.5, 1/3, 0.25, 0.3: `>` is strict
    and the CUDA kernel's fma-based union decides the last bit)."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(78)
    for n, grid in ((64, 4), (700, 8), (1500, 16), (5000, 32)):
        # integer boxes on a coarse lattice with a few sizes: many identical boxes, many pairs with small-integer
        # intersection / union ratios (e.g. 8x8 boxes shifted by 4 columns: inter 32, union 96 -> exactly 1/3)
        xy = rng.integers(0, grid, (n, 2)) * 4
        wh = rng.choice([4, 8, 12, 16], (n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = (rng.integers(0, 8, n) / 8.0).astype(np.float32)  # 8 distinct scores: long runs of exact ties
        b, sc = gpu(boxes), gpu(scores)
        for t in (0.5, 1.0 / 3.0, 0.25, 0.3, 0.0, 1.0):
            ref = tv.ops.nms(b, sc, t).cpu().numpy()
            got = ops.nms(b, sc, t, ops.IOU_TV_CUDA).cpu().numpy()
            assert np.array_equal(got, ref), (n, t, len(got), len(ref))
    # the pairs that sit exactly on a threshold exist in the data above (otherwise the test would prove nothing)
    a = np.array([[0, 0, 8, 8]], np.float32); c = np.array([[4, 0, 12, 8]], np.float32)
    inter = 4 * 8; union = 64 + 64 - inter
    assert inter / union == 1.0 / 3.0
    pair = gpu(np.concatenate([a, c])); s2 = gpu(np.array([0.9, 0.8], np.float32))
    for t in (1.0 / 3.0, np.nextafter(np.float32(1.0 / 3.0), np.float32(0)), np.nextafter(np.float32(1.0 / 3.0), np.float32(1))):
        assert np.array_equal(ops.nms(pair, s2, float(t), ops.IOU_TV_CUDA).cpu().numpy(), tv.ops.nms(pair, s2, float(t)).cpu().numpy()), t


def test_nms_batched_segments_with_ties_and_empty():
    """Many segments of different sizes (0, 1, ragged) in one call; tie order follows `tie`."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 5, 64, 200, 0, 777, 1500]
    cap = 1600
    rows = np.zeros((len(sizes), cap, 21), np.float32)
    tie = np.zeros((len(sizes), cap), np.int32)
    refs = []
    for gi, n in enumerate(sizes):
        p = synth.synth_pred(max(n, 1), 12, 30, 100 + gi, tie_decimals=1)[:n]
        cand, src = O.filter_xyxy(p, -1.0)  # everything passes; xyxy conversion
        perm = rng.permutation(n)         # scramble storage order like the atomic compaction does
        rows[gi, :n] = cand[perm]
        tie[gi, :n] = src[perm]
        keep = O.nms(cand[:, :4], cand[:, 4], 0.45, O.TV_CUDA)
        refs.append(cand[keep])
    count = torch.tensor(sizes, dtype=torch.int32, device=DEV)
    rows_d = gpu(rows)
    keep, n_keep = ops.nms_segments(rows_d, count, cap, 0.45, ops.IOU_TV_CUDA, tie=gpu(tie))
    out = ops.gather_rows(rows_d, keep, n_keep, 21, cap).cpu().numpy()
    nk = n_keep.cpu().numpy()
    for gi, n in enumerate(sizes):
        assert nk[gi] == len(refs[gi])
        assert np.array_equal(out[gi, :nk[gi]], refs[gi])


def test_nms_max_nms_cut():
    """utils/general.py:578-579: more than max_nms candidates -> only the best max_nms enter NMS."""
    n, cut = 3000, 1000
    p = synth.synth_pred(n, 40, 60, 9)
    cand, _ = O.filter_xyxy(p, -1.0)
    order = np.argsort(-cand[:, 4], kind="stable")[:cut]
    ref = order[O.nms(cand[order, :4], cand[order, 4], 0.5, O.TV_CUDA)]
    rows = gpu(cand[None])
    count = torch.tensor([n], dtype=torch.int32, device=DEV)
    keep, n_keep = ops.nms_segments(rows, count, n, 0.5, ops.IOU_TV_CUDA, max_nms=cut)
    assert np.array_equal(keep[0, :int(n_keep[0])].cpu().numpy(), ref)


def test_nms_large_segment_global_sort_path():
    """> 8192 candidates: sort runs out of the global scratch; dense crowd, heavy ties."""
    n = 20000
    p = synth.synth_pred(n, 300, 70, 21, extent=(1280.0, 1280.0), tie_decimals=2)
    cand, src = O.filter_xyxy(p, -1.0)
    ref = O.nms(cand[:, :4], cand[:, 4], 0.5, O.TV_CUDA)
    rows = gpu(cand[None])
    count = torch.tensor([n], dtype=torch.int32, device=DEV)
    keep, n_keep = ops.nms_segments(rows, count, n, 0.5, ops.IOU_TV_CUDA)
    assert int(n_keep[0]) == len(ref)
    assert np.array_equal(keep[0, :len(ref)].cpu().numpy(), ref)


# =============================================================================== rescale / merge
def test_rescale_golden(golden):
    g = golden("rescale")
    for i, (ih, iw, h0, w0) in enumerate(g["meta"]):
        b = g["c%d_boxes" % i]; k = g["c%d_kpts" % i]
        assert np.array_equal(ops.scale_coords_((ih, iw), gpu(b), (h0, w0)).cpu().numpy(), g["c%d_std_boxes" % i])
        assert np.array_equal(ops.scale_coords_((ih, iw), gpu(k), (h0, w0), kpt_step=3).cpu().numpy(),
                              g["c%d_std_kpts" % i])
        assert np.array_equal(ops.scale_coords_api_((ih, iw), gpu(b), (h0, w0)).cpu().numpy(), g["c%d_api_boxes" % i])
        assert np.array_equal(ops.scale_coords_api_((ih, iw), gpu(b), (h0, w0), True).cpu().numpy(),
                              g["c%d_api_round" % i])


def test_rescale_ratio_pad_golden(golden):
    """general.scale_coords with an explicit ratio_pad (utils/general.py:379-383; test.py:172), in place, on CUDA and
    CPU tensors, against the reference's outputs."""
    from face_detection_multi_scale_b200 import general
    g = golden("rescale_rp")
    for i, (ih, iw, h0, w0) in enumerate(g["meta"]):
        for j in range(3):
            gain, px, py = (float(v) for v in g["c%d_%d_rp" % (i, j)])
            rp = ((gain, gain), (px, py)) if j < 2 else (gain, (px, py))
            b = gpu(g["c%d_%d_boxes" % (i, j)]); k = torch.from_numpy(g["c%d_%d_kpts" % (i, j)].copy())
            r = general.scale_coords((ih, iw), b, (h0, w0), rp)
            assert r is b and np.array_equal(b.cpu().numpy(), g["c%d_%d_std_boxes" % (i, j)])
            general.scale_coords((ih, iw), k, (h0, w0), rp, kpt_label=5, step=3)
            assert np.array_equal(k.numpy(), g["c%d_%d_std_kpts" % (i, j)])


def test_weighted_nms_golden(golden):
    g = golden("merge")
    for ci in range(5):
        dets = g["m%d_dets" % ci]
        n_scales, iou = g["m%d_meta" % ci]
        rows = ops.weighted_rows(gpu(dets), int(n_scales))
        w = rows[0, :, 4].cpu().numpy().astype(np.float64) / dets[:, 4]
        assert np.allclose(w, g["m%d_weights" % ci], rtol=1e-6)
        count = torch.tensor([dets.shape[0]], dtype=torch.int32, device=DEV)
        keep, n_keep = ops.nms_segments(rows, count, dets.shape[0], float(iou), ops.IOU_TV_CPU, max_nms=0)
        k = keep[0, :int(n_keep[0])].cpu().numpy()
        assert np.array_equal(dets[k], g["m%d_out" % ci])


# =============================================================================== end to end
@pytest.mark.parametrize("case", H.E2E_CASES, ids=[c[0] for c in H.E2E_CASES])
def test_e2e_postprocess_golden(golden, case):
    """Raw heads of every scale -> fdms_postprocess (one C call) == the reference's detect_multi_scale."""
    g = golden("e2e")
    name, h, w, sizes, n_faces, conf, iou, api, seed = case
    heads, shapes = H.e2e_heads(h, w, sizes, n_faces, api, seed)
    sl = [[gpu(l[None]) for l in heads[s]] for s in shapes]
    plan = ops.PostPlan(sl, synth.P5_STRIDES, synth.P5_ANCHORS, [s[0] for s in shapes], (h, w), conf, iou, cap=4096,
                        out_cap=4096, iou_mode=ops.IOU_TV_CPU)  # the golden run was on CPU
    out = plan.run()
    torch.cuda.synchronize()
    n = int(out.n_out[0])
    ref = g[name + "_final"]
    got = out.dets[0, :n].cpu().numpy().astype(np.float64)
    assert got.shape[0] == ref.shape[0]
    if ref.size:
        assert np.array_equal(got, ref)
    # per-scale survivors as well
    nk = plan.cands.n_keep.cpu().numpy()
    for si in range(len(sizes)):
        assert nk[si] == g["%s_scale%d" % (name, si)].shape[0]


# =============================================================================== edge cases
def test_edge_thresholds_and_degenerate_boxes():
    """conf/iou thresholds at the ends of their range, NaN / zero-area / identical boxes, negative coordinates."""
    rng = np.random.default_rng(3)
    n = 700
    xy = rng.uniform(-50, 300, (n, 2)); wh = rng.uniform(0, 60, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    boxes[::50, 2:] = boxes[::50, :2]            # zero-area boxes (IoU 0/0 = NaN never suppresses)
    boxes[5] = boxes[6] = boxes[7]               # exact duplicates
    boxes[11, 0] = np.nan                        # NaN coordinate: never suppressed, never suppresses
    scores = rng.uniform(0, 1, n).astype(np.float32)
    scores[5] = scores[6] = scores[7] = 0.5
    for t in (0.0, 1e-9, 0.3, 0.999999, 1.0):
        for mode in (ops.IOU_TV_CPU, ops.IOU_TV_CUDA):
            ref = O.nms(boxes, scores, t, mode)
            got = ops.nms(gpu(boxes), gpu(scores), t, mode).cpu().numpy()
            assert np.array_equal(got, ref), (t, mode)


@pytest.mark.parametrize("conf", [0.0, 1.0, 0.999999])
def test_decode_filter_extreme_conf_thresholds(conf):
    levels = _crowd("crowdB")
    ref_pred = O.decode_head(levels, synth.P5_STRIDES, synth.P5_ANCHORS)
    c = ops.decode_filter([[gpu(l) for l in levels]], synth.P5_STRIDES, synth.P5_ANCHORS, conf, cap=ref_pred.shape[1])
    for b in range(ref_pred.shape[0]):
        rc, rr = O.filter_xyxy(ref_pred[b], conf)
        cand, rows = _sorted_cands(c, b)
        assert np.array_equal(rows, rr)
        if len(rr):
            assert np.max(np.abs(cand - rc)) <= TOL_PX


def test_postprocess_all_empty_and_single_scale():
    """No candidate anywhere (background only) and a 1-scale detector (weights: S=1 => s >= S-2 and s == S//2)."""
    rng = np.random.default_rng(0)
    lv = [rng.standard_normal((2, 63, 12, 20)).astype(np.float32) - 8.0, rng.standard_normal((2, 63, 6, 10)).astype(np.float32) - 8.0,
          rng.standard_normal((2, 63, 3, 5)).astype(np.float32) - 8.0]
    plan = ops.PostPlan([[gpu(l) for l in lv]], synth.P5_STRIDES, synth.P5_ANCHORS, [96], (270, 480), 0.5, 0.5, cap=64,
                        out_cap=64)
    out = plan.run()
    assert out.n_out.cpu().tolist() == [0, 0] and plan.cands.count.cpu().tolist() == [0, 0]
    heads, shapes = H.e2e_heads(200, 200, [128], 10, False, 34)
    sl = [[gpu(l[None]) for l in heads[shapes[0]]]]
    plan = ops.PostPlan(sl, synth.P5_STRIDES, synth.P5_ANCHORS, [shapes[0][0]], (200, 200), 0.5, 0.5, cap=512, out_cap=512,
                        iou_mode=ops.IOU_TV_CPU)
    out = plan.run()
    ref, _ = O.detect_from_heads([[l[None] for l in heads[shapes[0]]]], shapes, (200, 200, 3), synth.P5_STRIDES,
                                 synth.P5_ANCHORS, 0.5, 0.5, O.TV_CPU)
    n = int(out.n_out[0])
    assert n == len(ref) and np.array_equal(out.dets[0, :n].cpu().numpy().astype(np.float64), ref)


def test_w6_maximum_rows_single_segment():
    """The largest pred of the reference's configs: yolov7-w6 at 1536 (85 680 rows), every row a candidate, so the
    candidate count exceeds max_nms = 30000 and the best 30000 enter NMS (utils/general.py:578-579)."""
    rows = 85680
    p = synth.synth_pred(rows, 2000, 40, 5, extent=(896.0, 1536.0))
    p[:, 4] = np.maximum(p[:, 4], 0.8); p[:, 5] = np.maximum(p[:, 5], 0.8)
    from face_detection_multi_scale_b200 import general
    res = general.non_max_suppression(torch.from_numpy(p[None]).to(DEV), 0.5, 0.45, iou_mode=ops.IOU_TV_CPU)[0]
    ref = O.non_max_suppression(p[None], 0.5, 0.45, O.TV_CPU)[0]
    assert res.shape == ref.shape and np.array_equal(res.cpu().numpy(), ref)


def test_non_finite_head_values_do_not_hang_and_match_oracle():
    """NaN / +-Inf logits in the raw head (a diverged backbone): NaN confidences never pass the filter, infinite
    logits saturate the sigmoid; the candidate sets still equal the oracle's and the NMS terminates (IoU with
    non-finite boxes is NaN, which never suppresses)."""
    levels = [l.copy() for l in _crowd("crowdB")]
    rng = np.random.default_rng(0)
    for l in levels:
        flat = l.reshape(-1)
        idx = rng.choice(flat.size, 40, replace=False)
        flat[idx[:15]] = np.nan
        flat[idx[15:28]] = np.inf
        flat[idx[28:]] = -np.inf
    c = ops.decode_filter([[gpu(l) for l in levels]], synth.P5_STRIDES, synth.P5_ANCHORS, 0.3, cap=4096)
    torch.cuda.synchronize()
    with np.errstate(all="ignore"):
        pred = O.decode_head(levels, synth.P5_STRIDES, synth.P5_ANCHORS)
    for b in range(2):
        with np.errstate(all="ignore"):
            rc, rr = O.filter_xyxy(pred[b], 0.3)
        cand, rows = _sorted_cands(c, b)
        assert np.array_equal(rows, rr)
        assert np.array_equal(np.isnan(cand), np.isnan(rc))
        fin = np.isfinite(rc)
        assert np.max(np.abs(cand[fin] - rc[fin]), initial=0) <= TOL_PX
    keep, n_keep = ops.nms_segments(c.cand, c.count, c.cap, 0.45, ops.IOU_TV_CPU, tie=c.row)
    torch.cuda.synchronize()
    for b in range(2):
        with np.errstate(all="ignore"):
            rc, rr = O.filter_xyxy(pred[b], 0.3)
            ref = O.nms(rc[:, :4], rc[:, 4], 0.45, O.TV_CPU)
        cand, rows = _sorted_cands(c, b)
        got_rows = c.row[b][keep[b, :int(n_keep[b])].long()].cpu().numpy()
        assert np.array_equal(got_rows, rr[ref])


@pytest.mark.parametrize("mode", [ops.IOU_TV_CPU, ops.IOU_TV_CUDA, ops.IOU_NCNN])
def test_nms_fuzz_vs_oracle(mode):
    """Randomised differential test of the batched NMS against the oracle: segment sizes straddling every kernel
    boundary (32-key runs, 256-candidate steps, the 1024-box fused kernel, sorted path), tight and loose clusters,
    quantised scores (exact ties), integer corners (IoU exactly on the threshold), several segments per call."""
    rng = np.random.default_rng(1234 + mode)
    sizes = [1, 2, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 1500, 2600]
    for it in range(48):
        n_seg = int(rng.integers(1, 5))
        cap = int(max(rng.choice(sizes, n_seg)))
        rows = np.zeros((n_seg, cap, 6), np.float32)
        counts = np.zeros((n_seg,), np.int32)
        thr = float(rng.choice([0.1, 0.3, 0.45, 0.5, 1 / 3, 0.7]))
        for g in range(n_seg):
            n = int(rng.integers(1, cap + 1)) if g else cap
            k = int(rng.integers(1, max(2, n // int(rng.integers(2, 12)) + 1)))
            ctr = rng.uniform(0, 600, (k, 2))
            sz = rng.uniform(6, 120, (k, 2))
            idx = rng.integers(0, k, n)
            c = ctr[idx] + rng.normal(0, rng.choice([0.02, 0.1, 0.4]), (n, 2)) * sz[idx]
            wh = sz[idx] * rng.uniform(0.7, 1.4, (n, 2))
            b = np.concatenate([c - wh / 2, c + wh / 2], 1)
            if it % 3 == 0:
                b = np.round(b)
            rows[g, :n, :4] = b
            rows[g, :n, 4] = np.round(rng.uniform(0.05, 1, n), int(rng.choice([2, 3, 6])))
            counts[g] = n
        ops.debug_poison_smem(0xFF if it & 1 else 0x7F)
        keep, n_keep = ops.nms_segments(gpu(rows), gpu(counts), cap, thr, mode, max_nms=0)
        torch.cuda.synchronize()
        for g in range(n_seg):
            n = counts[g]
            ref = O.nms(rows[g, :n, :4], rows[g, :n, 4], thr, mode)
            got = keep[g, :int(n_keep[g])].cpu().numpy()
            assert np.array_equal(got, ref), (it, g, n, thr, mode)


@pytest.mark.parametrize("seed", range(3))
def test_letterbox_fuzz_vs_oracle(seed):
    """Randomised letterbox pyramid against the oracle: random source sizes (odd widths -> rows off the 16-byte grid,
    portrait and landscape), random target sizes incl. exact 1:1 / 2:1 / 3:1 / 3:2 factors, api mode and plain mode,
    two or three scales per launch, fp16 CHW with the channel flip -- every tile kind and staging path, bit for bit."""
    rng = np.random.default_rng(99 if seed == 0 else 1000 + seed)
    for it in range(40):
        if it % 4 == 0:  # shapes that hit the exact factors
            base = int(rng.choice([32, 64, 96]))
            k = rng.choice([1, 2, 3, 1.5])
            h, w = int(base * k * rng.choice([1, 2])), int(base * k * rng.choice([2, 3]))
            if rng.random() < 0.5:
                h, w = w, h
            sizes = [int(max(h, w) / k), int(rng.integers(24, 200))]
        else:
            h, w = int(rng.integers(9, 300)), int(rng.integers(9, 300))
            sizes = [int(s) for s in rng.integers(16, 260, int(rng.integers(2, 4)))]
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        api = it % 3 == 0
        if api:
            s_ = max(h, w)
            geoms = [_lib.letterbox_geometry(s_, s_, size, auto=False, stride=32) for size in sizes]
            outs = _pyr(img, geoms, torch.float16, layout="chw", flip=False, virt=(s_, s_))
            refs = [(O.preprocess_api(img, size).astype(np.float32) / np.float32(255)).astype(np.float16) for size in sizes]
        else:
            geoms = [_lib.letterbox_geometry(h, w, size, auto=True, stride=32) for size in sizes]
            outs = _pyr(img, geoms, torch.float16)
            refs = [O.pack_chw(O.letterbox(img, size, auto=True, stride=32)[0], True, np.float16) for size in sizes]
        for size, got, ref in zip(sizes, outs, refs):
            assert got.shape == ref.shape, (it, h, w, size)
            assert np.array_equal(got.view(np.uint16), ref.view(np.uint16)), (it, h, w, size, api)


@pytest.mark.parametrize("seed", range(2))
def test_postprocess_fuzz_vs_oracle(seed):
    """Randomised whole post-processing (decode + filter -> per-scale NMS -> rescale -> cross-scale merge, ONE
    fdms_postprocess call) against the oracle's detect_from_heads: P5 and P6 heads, 1-3 scales of random (tiny, odd
    nx / ny) map sizes, batch 1-3, fp16 / fp32 maps, random thresholds, both torchvision IoU flavours."""
    rng = np.random.default_rng(500 + seed)
    for it in range(12):
        p6 = bool(rng.integers(0, 2))
        strides, anchors = (synth.P6_STRIDES, synth.P6_ANCHORS) if p6 else (synth.P5_STRIDES, synth.P5_ANCHORS)
        smax = strides[-1]
        n_scales, bs = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        dtype = np.float16 if rng.integers(0, 2) else np.float32
        shapes = [(int(rng.integers(1, 6)) * smax, int(rng.integers(1, 7)) * smax) for _ in range(n_scales)]
        conf = float(rng.choice([0.05, 0.25, 0.5, 0.7]))
        iou = float(rng.choice([0.3, 0.45, 0.5]))
        mode = ops.IOU_TV_CPU if rng.integers(0, 2) else ops.IOU_TV_CUDA
        h0, w0 = int(rng.integers(40, 2000)), int(rng.integers(40, 2000))
        bias = float(rng.choice([-3.0, -1.5, 0.0]))
        sl = []
        for ih, iw in shapes:
            lv = []
            for s in strides:
                a = (rng.standard_normal((bs, 63, ih // s, iw // s)) * 1.5).astype(np.float32)
                a[:, 4::21] += bias  # objectness logits of the three anchors
                lv.append(a.astype(dtype))
            sl.append(lv)
        rows_max = max(sum(3 * (ih // s) * (iw // s) for s in strides) for ih, iw in shapes)
        ops.debug_poison_smem(0xFF if it & 1 else 0x7F)
        plan = ops.PostPlan([[gpu(l) for l in lv] for lv in sl], strides, anchors, [s[0] for s in shapes], (h0, w0),
                            conf, iou, cap=rows_max, out_cap=rows_max * n_scales, iou_mode=mode)
        out = plan.run()
        torch.cuda.synchronize()
        counts = plan.cands.count.cpu().numpy()
        nk = plan.cands.n_keep.cpu().numpy()
        for b in range(bs):
            heads_b = [[l[b:b + 1] for l in lv] for lv in sl]
            ref, per_scale = O.detect_from_heads(heads_b, shapes, (h0, w0, 3), strides, anchors, conf, iou, mode)
            tag = (seed, it, b, p6, shapes, conf, iou, mode, str(dtype))
            for si in range(n_scales):
                pred = O.decode_head(heads_b[si], strides, anchors)
                assert counts[b * n_scales + si] == len(O.filter_xyxy(pred[0], conf)[1]), tag
                assert nk[b * n_scales + si] == per_scale[si].shape[0], tag
            n = int(out.n_out[b])
            got = out.dets[b, :n].cpu().numpy().astype(np.float64)
            assert got.shape[0] == (ref.shape[0] if ref.size else 0), tag
            if ref.size:
                assert np.array_equal(got, ref), tag


@pytest.mark.parametrize("seed", range(2))
def test_letterbox_fuzz_layouts_batches_and_pitches(seed):
    """The letterbox fuzz over the remaining degrees of freedom: output type (u8 / fp16 / fp32), layout (CHW / HWC),
    channel flip, batches of 1-3 images, sources that are views with a padded row pitch (aligned and not), border
    values other than 114 -- bit for bit against the oracle."""
    rng = np.random.default_rng(700 + seed)
    for it in range(30):
        h, w = int(rng.integers(9, 260)), int(rng.integers(9, 260))
        if it % 5 == 0:  # exact-factor shapes
            k = float(rng.choice([1, 2, 3, 1.5]))
            base = int(rng.choice([32, 64]))
            h, w = int(base * k), int(base * k * rng.choice([1, 2]))
            sizes = [int(max(h, w) / k), int(rng.integers(24, 160))]
        else:
            sizes = [int(s) for s in rng.integers(16, 200, int(rng.integers(1, 4)))]
        bs = int(rng.integers(1, 4))
        pad = int(rng.choice([0, 0, 1, 4, 5, 16]))
        dtype = [torch.uint8, torch.float16, torch.float32][int(rng.integers(0, 3))]
        layout = "chw" if rng.integers(0, 2) else "hwc"
        flip = bool(rng.integers(0, 2))
        border = int(rng.choice([114, 0, 255, 7]))
        api = it % 3 == 1
        full = rng.integers(0, 256, (bs, h, w + pad, 3), dtype=np.uint8)
        imgs = np.ascontiguousarray(full[:, :, :w])
        src = gpu(full)[:, :, :w, :]
        ops.debug_poison_smem(0xFF if it & 1 else 0x7F)
        if api:
            s_ = max(h, w)
            geoms = [_lib.letterbox_geometry(s_, s_, size, auto=False, stride=32) for size in sizes]
            outs = ops.pyramid(src, geoms, out_dtype=dtype, layout=layout, flip_rb=flip, border=border, virt_hw=(s_, s_))
        else:
            geoms = [_lib.letterbox_geometry(h, w, size, auto=bool(it & 1), stride=32) for size in sizes]
            outs = ops.pyramid(src, geoms, out_dtype=dtype, layout=layout, flip_rb=flip, border=border)
        torch.cuda.synchronize()
        for size, got in zip(sizes, outs):
            got = got.cpu().numpy()
            for b in range(bs):
                if api:
                    u8 = O.letterbox(imgs[b], size, color=(border,) * 3, auto=False, stride=32, virt_hw=(s_, s_))[0]
                else:
                    u8 = O.letterbox(imgs[b], size, color=(border,) * 3, auto=bool(it & 1), stride=32)[0]
                if dtype == torch.uint8:
                    ref = u8[:, :, ::-1] if flip else u8
                    ref = ref.transpose(2, 0, 1) if layout == "chw" else ref
                else:
                    ref = O.pack_chw(u8, flip, np.float16 if dtype == torch.float16 else np.float32)
                    ref = ref if layout == "chw" else ref.transpose(1, 2, 0)
                tag = (seed, it, b, h, w, pad, size, str(dtype), layout, flip, api)
                assert got[b].shape == ref.shape, tag
                assert np.array_equal(got[b], ref), tag


def test_letterbox_fuzz_multi_tile_shapes():
    """The small fuzz shapes fit one 320-column tile; these do not: sources and targets of several hundred to ~1700
    pixels (several tile columns, partial last tiles, upscales past one tile, odd pitches), fp16 CHW with the flip --
    the detector's configuration -- bit for bit against the oracle."""
    rng = np.random.default_rng(808)
    for it in range(14):
        h, w = int(rng.integers(200, 1300)), int(rng.integers(330, 1700))
        if it % 4 == 3:
            h, w = w, h
        if it % 7 == 6:  # an exact 3:2 / 2:1 pyramid next to a general scale
            h, w = 360, 960
            sizes = [640, 480, int(rng.integers(330, 900))]
        else:
            sizes = [int(s) for s in rng.integers(330, 1200, 2)]
        pad = int(rng.choice([0, 1, 4, 16]))
        api = it % 3 == 2
        full = rng.integers(0, 256, (1, h, w + pad, 3), dtype=np.uint8)
        img = np.ascontiguousarray(full[0, :, :w])
        src = gpu(full)[:, :, :w, :]
        ops.debug_poison_smem(0xFF if it & 1 else 0x7F)
        if api:
            s_ = max(h, w)
            geoms = [_lib.letterbox_geometry(s_, s_, size, auto=False, stride=32) for size in sizes]
            outs = ops.pyramid(src, geoms, out_dtype=torch.float16, layout="chw", flip_rb=False, virt_hw=(s_, s_))
            refs = [(O.preprocess_api(img, size).astype(np.float32) / np.float32(255)).astype(np.float16) for size in sizes]
        else:
            geoms = [_lib.letterbox_geometry(h, w, size, auto=True, stride=32) for size in sizes]
            outs = ops.pyramid(src, geoms, out_dtype=torch.float16)
            refs = [O.pack_chw(O.letterbox(img, size, auto=True, stride=32)[0], True, np.float16) for size in sizes]
        torch.cuda.synchronize()
        for size, got, ref in zip(sizes, outs, refs):
            got = got[0].cpu().numpy()
            assert got.shape == ref.shape, (it, h, w, pad, size, api)
            assert np.array_equal(got.view(np.uint16), ref.view(np.uint16)), (it, h, w, pad, size, api)
