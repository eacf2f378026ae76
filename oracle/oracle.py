"""ctypes/numpy front-end of the CPU parity oracle (oracle/fdms_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Each function cites the reference
file:line it restates (paths relative to /root/reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Geom(C.Structure):
    _fields_ = [("dst_h", C.c_int), ("dst_w", C.c_int), ("unpad_h", C.c_int), ("unpad_w", C.c_int),
                ("top", C.c_int), ("bottom", C.c_int), ("left", C.c_int), ("right", C.c_int),
                ("resize", C.c_int), ("ratio_w", C.c_double), ("ratio_h", C.c_double),
                ("dw", C.c_double), ("dh", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libfdms_oracle.so")
    src = os.path.join(_HERE, "fdms_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libfdms_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.fo_nms.restype = C.c_int64
        _LIB.fo_filter_xyxy.restype = C.c_int64
        _LIB.fo_weighted_nms.restype = C.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- B1 geometry: utils/datasets.py:873-903 ---------------------------------------------------
def letterbox_geometry(src_h, src_w, new_shape=(640, 640), auto=True, scaleFill=False, scaleup=True, stride=32):
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    g = Geom()
    lib().fo_letterbox_geometry(int(src_h), int(src_w), int(new_shape[0]), int(new_shape[1]), int(auto),
                                int(scaleFill), int(scaleup), int(stride), C.byref(g))
    return g


# ---- B2 cv2.resize(INTER_LINEAR) u8 -----------------------------------------------------------
def resize_linear_u8(img, dst_h, dst_w, virt_hw=None):
    img = np.ascontiguousarray(img)
    h, w, cn = img.shape
    vh, vw = virt_hw if virt_hw else (h, w)
    out = np.empty((dst_h, dst_w, cn), np.uint8)
    rc = lib().fo_resize_linear_u8(_p(img), h, w, C.c_int64(w * cn), vh, vw, _p(out), dst_h, dst_w,
                                   C.c_int64(dst_w * cn), cn)
    assert rc == 0
    return out


# ---- letterbox: utils/datasets.py:873-903 -----------------------------------------------------
def letterbox(img, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32,
              virt_hw=None):
    """Returns (img, (rw, rh), (dw, dh)) like the reference.  virt_hw: resize over a zero-padded
    virtual source of that size (the api path, utils/preprocess_yolo_predict.py:273-290,370)."""
    img = np.ascontiguousarray(img)
    h, w, _ = img.shape
    vh, vw = virt_hw if virt_hw else (h, w)
    g = letterbox_geometry(vh, vw, new_shape, auto, scaleFill, scaleup, stride)
    out = np.empty((g.dst_h, g.dst_w, 3), np.uint8)
    rc = lib().fo_letterbox_u8(_p(img), h, w, C.c_int64(w * 3), vh, vw, C.byref(g), int(color[0]), _p(out))
    assert rc == 0
    if not (color[0] == color[1] == color[2]):  # cv2.copyMakeBorder(value=color): one value per channel (:902)
        inner = np.zeros((g.dst_h, g.dst_w), bool)
        inner[g.top:g.top + g.unpad_h, g.left:g.left + g.unpad_w] = True
        for c in (1, 2):
            out[..., c][~inner] = int(color[c])
    return out, (g.ratio_w, g.ratio_h), (g.dw, g.dh)


def preprocess_api(img_rgb, img_size, stride=32):
    """utils/preprocess_yolo_predict.py:345-378 from an already-decoded RGB array: pad to square
    (top-left, zeros) -> letterbox(auto=False) -> CHW.  Returns (3, size, size) uint8."""
    h, w, _ = img_rgb.shape
    s = max(h, w)
    out, _, _ = letterbox(img_rgb, img_size, auto=False, stride=stride, virt_hw=(s, s))
    return np.ascontiguousarray(out.transpose(2, 0, 1))


# ---- B3 pack: multi_scale_face_detector.py:96-105 ----------------------------------------------
def pack_chw(img_hwc, flip=True, dtype=np.float16):
    img_hwc = np.ascontiguousarray(img_hwc)
    h, w, _ = img_hwc.shape
    if dtype == np.float16:
        out = np.empty((3, h, w), np.uint16)
        lib().fo_pack_chw(_p(img_hwc), h, w, int(flip), 1, _p(out))
        return out.view(np.float16)
    out = np.empty((3, h, w), np.float32)
    lib().fo_pack_chw(_p(img_hwc), h, w, int(flip), 0, _p(out))
    return out


# ---- B5 decode: models/yolo.py:267-308 ---------------------------------------------------------
def decode_head(levels, strides, anchors_px, no=21, na=3):
    """levels: list of (bs, na*no, ny, nx) arrays (fp16 or fp32; up-cast to fp32 first).
    anchors_px: (nl, na, 2) pixels.  Returns pred (bs, rows, no) fp32, rows level-major."""
    levels = [np.ascontiguousarray(np.asarray(l, dtype=np.float32)) for l in levels]
    bs = levels[0].shape[0]
    rows = sum(na * l.shape[2] * l.shape[3] for l in levels)
    pred = np.empty((bs, rows, no), np.float32)
    anchors_px = np.asarray(anchors_px, np.float32).reshape(len(levels), na, 2)
    off = 0
    for i, l in enumerate(levels):
        _, ch, ny, nx = l.shape
        assert ch == na * no
        a = np.ascontiguousarray(anchors_px[i])
        lib().fo_decode_level(_p(l), bs, na, no, ny, nx, C.c_float(float(strides[i])), _p(a), C.c_int64(rows),
                              C.c_int64(off), _p(pred))
        off += na * ny * nx
    return pred


def decode_head_general(levels, strides, anchors_px, na, no, nkpt):
    """Any head variant of models/yolo.py (Detect :59-106, IDetect :152-207, IKeypoint :267-308): levels = list of
    (bs, na*no, ny, nx); nkpt > 0 -> keypoint head (no == 6 + 3*nkpt), nkpt == 0 -> plain head (no == 5 + nc)."""
    levels = [np.ascontiguousarray(np.asarray(l, dtype=np.float32)) for l in levels]
    assert nkpt == 0 or no == 6 + 3 * nkpt
    bs = levels[0].shape[0]
    rows = sum(na * l.shape[2] * l.shape[3] for l in levels)
    pred = np.empty((bs, rows, no), np.float32)
    anchors_px = np.asarray(anchors_px, np.float32).reshape(len(levels), na, 2)
    off = 0
    for i, l in enumerate(levels):
        _, ch, ny, nx = l.shape
        assert ch == na * no
        a = np.ascontiguousarray(anchors_px[i])
        lib().fo_decode_level_head(_p(l), bs, na, no, nkpt, ny, nx, C.c_float(float(strides[i])), _p(a), C.c_int64(rows),
                                   C.c_int64(off), _p(pred))
        off += na * ny * nx
    return pred


def decode_head_half(levels, strides, anchors_px, na=3, no=21, nkpt=5):
    """The same decode as ATen computes it for a HALF model (the reference's default on a GPU, msfd:49 `model.half()`):
    the conv outputs, the sigmoid, `* 2.`, `- 0.5`, `** 2`, `* anchor_grid` are half tensors (each operation computes in
    float32 and rounds to half), `self.grid[i]` is float32 (models/yolo.py:310-313), so `+ grid` and `* stride` are float32;
    the keypoint coordinates are stored back into the half tensor `x_kpt` (:293-295) and `torch.cat` (:297) promotes the
    row: pred is FLOAT32 with half-valued columns 2.. and float32 xy.  levels: list of (bs, na*no, ny, nx) float16 arrays.
    Pinned op by op against the reference modules on half tensors (tests/ref_live_check.py); numpy, small cases."""
    f32, f16 = np.float32, np.float16
    h = lambda a: np.asarray(a, f32).astype(f16).astype(f32)  # noqa: E731  one half rounding
    sig = lambda a: h(f32(1.0) / (f32(1.0) + np.exp(-np.asarray(a, f32))))  # noqa: E731
    anchors_px = np.asarray(anchors_px, f32).reshape(len(levels), na, 2)
    n_sig = 6 if nkpt > 0 else no
    outs = []
    for i, l in enumerate(levels):
        l = np.asarray(l)
        assert l.dtype == f16, "decode_head_half takes the half conv outputs"
        bs, ch, ny, nx = l.shape
        assert ch == na * no
        v = l.reshape(bs, na, no, ny, nx).transpose(0, 1, 3, 4, 2).astype(f32)
        gy, gx = np.meshgrid(np.arange(ny, dtype=f32), np.arange(nx, dtype=f32), indexing="ij")
        st = f32(strides[i])
        o = np.empty_like(v)
        s = sig(v[..., :n_sig])
        o[..., 0] = (h(h(s[..., 0] * f32(2)) - f32(0.5)) + gx) * st
        o[..., 1] = (h(h(s[..., 1] * f32(2)) - f32(0.5)) + gy) * st
        anc = h(anchors_px[i])  # the anchor_grid buffer of a half model
        for c in (2, 3):
            t = h(s[..., c] * f32(2))
            o[..., c] = h(h(t * t) * anc[:, c - 2][None, :, None, None])
        o[..., 4:n_sig] = s[..., 4:]
        for j in range(nkpt):
            for k, g in ((0, gx), (1, gy)):
                t = h(h(v[..., 6 + 3 * j + k] * f32(2)) - f32(0.5))
                o[..., 6 + 3 * j + k] = h((t + g) * st)
            o[..., 8 + 3 * j] = sig(v[..., 8 + 3 * j])
        outs.append(o.reshape(bs, -1, no))
    return np.concatenate(outs, 1)


# ---- B6 filter: utils/general.py:515-563 -------------------------------------------------------
def filter_xyxy(pred_img, conf_thres):
    """pred_img: (rows, 21) fp32 for ONE image.  Returns (cand (n,21), src_row (n,))."""
    pred_img = np.ascontiguousarray(pred_img, np.float32)
    rows, no = pred_img.shape
    cand = np.empty((rows, no), np.float32)
    src = np.empty((rows,), np.int32)
    n = lib().fo_filter_xyxy(_p(pred_img), C.c_int64(rows), no, C.c_float(np.float32(conf_thres)), _p(cand), _p(src))
    return cand[:n].copy(), src[:n].copy()


# ---- B7 NMS: torchvision.ops.nms ---------------------------------------------------------------
TV_CPU, TV_CUDA, NCNN = 0, 1, 2  # NCNN: cpp/yolov7-face-ncnn/src/yolov7face.cpp:43-74 (+1 extents, >=)


def nms(boxes, scores, iou_thres, mode=TV_CPU):
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, np.float32)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), np.int64)
    k = lib().fo_nms(_p(boxes), _p(scores), C.c_int64(n), C.c_double(float(iou_thres)), int(mode), _p(keep))
    return keep[:k].copy()


MAX_NMS = 30000  # utils/general.py:520


def non_max_suppression(pred, conf_thres=0.25, iou_thres=0.45, mode=TV_CPU, labels=()):
    """utils/general.py:506-600 with kpt_label=5, nc=1, classes=None.  pred (bs, rows, 21) fp32.
    Returns list of (n,21) arrays; empty images give (0,6) zeros like the reference (:527).
    labels (autolabelling, :533-541; 6-column predictions only): per image (n, 5) [cls, x, y, w, h], appended after
    the obj > thres filter with obj = cls = 1."""
    out = []
    for xi in range(pred.shape[0]):
        cand, _ = filter_xyxy(pred[xi], conf_thres)
        if labels and len(labels[xi]) and 1.0 > np.float32(conf_thres):
            lab = np.asarray(labels[xi], np.float32).reshape(-1, 5)
            v = np.zeros((len(lab), cand.shape[1]), np.float32)
            half_w, half_h = lab[:, 3] / np.float32(2), lab[:, 4] / np.float32(2)
            v[:, 0], v[:, 1], v[:, 2], v[:, 3] = lab[:, 1] - half_w, lab[:, 2] - half_h, lab[:, 1] + half_w, lab[:, 2] + half_h
            v[:, 4] = 1.0
            cand = np.concatenate([cand, v], 0)
        if cand.shape[0] == 0:
            out.append(np.zeros((0, 6), np.float32))
            continue
        if cand.shape[0] > MAX_NMS:  # :578-579 (the reference's argsort is not stable; ours is)
            cand = cand[np.argsort(-cand[:, 4], kind="stable")[:MAX_NMS]]
        keep = nms(cand[:, :4], cand[:, 4], iou_thres, mode)
        out.append(cand[keep])
    return out


# ---- B8 rescale --------------------------------------------------------------------------------
def scale_coords_api(in_hw, boxes, img0_shape, do_round=False):
    """utils/preprocess_yolo_predict.py:122-157; in place on an (n,4) fp32 array."""
    assert boxes.dtype == np.float32 and boxes.flags.c_contiguous
    lib().fo_scale_coords_api(_p(boxes), C.c_int64(boxes.shape[0]), int(in_hw[0]), int(img0_shape[0]),
                              int(img0_shape[1]), int(do_round))
    return boxes


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None, kpt_label=False, step=2):
    """utils/general.py:374-404; in place on a contiguous (n,cols) fp32 array."""
    assert coords.dtype == np.float32 and coords.flags.c_contiguous
    n, cols = coords.shape
    if ratio_pad is not None:  # :379-383
        gain = ratio_pad[0][0] if isinstance(ratio_pad[0], (list, tuple)) else ratio_pad[0]
        lib().fo_scale_coords_rp(_p(coords), C.c_int64(n), cols, C.c_double(float(gain)),
                                 C.c_double(float(ratio_pad[1][0])), C.c_double(float(ratio_pad[1][1])),
                                 int(img0_shape[0]), int(img0_shape[1]), int(step) if kpt_label else 0)
        return coords
    lib().fo_scale_coords_std(_p(coords), C.c_int64(n), cols, int(img1_shape[0]), int(img1_shape[1]),
                              int(img0_shape[0]), int(img0_shape[1]), int(step) if kpt_label else 0)
    return coords


# ---- B9 merge: multi_scale_face_detector.py:168-240 ---------------------------------------------
def scale_weights(dets, n_scales):
    dets = np.ascontiguousarray(dets, np.float64).reshape(-1, 7)
    w = np.empty((dets.shape[0],), np.float64)
    lib().fo_scale_weights(_p(dets), C.c_int64(dets.shape[0]), int(n_scales), _p(w))
    return w


def weighted_nms(dets, n_scales, iou_thres):
    dets = np.ascontiguousarray(dets, np.float64).reshape(-1, 7)
    n = dets.shape[0]
    if n == 0:
        return dets
    keep = np.empty((n,), np.int64)
    k = lib().fo_weighted_nms(_p(dets), C.c_int64(n), int(n_scales), C.c_double(float(iou_thres)), _p(keep))
    return dets[keep[:k]]


def detect_from_heads(heads_per_scale, in_hw_per_scale, img0_shape, strides, anchors_px, conf_thres, iou_thres,
                      mode=TV_CPU, half_model=False):
    """multi_scale_face_detector.py:109-166,242-288 for ONE image, starting from the raw head
    tensors of each scale (the backbone is outside the hot path).  heads_per_scale[s] is a list of
    (1, 63, ny, nx) arrays.  Returns ((k,7) float64, per-scale (n,7) list).
    half_model: the levels are the float16 conv outputs of a half model (msfd:48-49) and the head decode rounds where
    ATen's half tensors round (decode_head_half); pred is float32 and everything after it is unchanged, as in the reference."""
    per_scale = []
    for s, levels in enumerate(heads_per_scale):
        pred = decode_head_half(levels, strides, anchors_px) if half_model else decode_head(levels, strides, anchors_px)
        det = non_max_suppression(pred, conf_thres, iou_thres, mode)[0]
        if det.shape[0] == 0:
            per_scale.append(np.zeros((0, 7)))
            continue
        boxes = np.ascontiguousarray(det[:, :4])
        scale_coords_api(in_hw_per_scale[s], boxes, img0_shape, do_round=True)  # msfd:144
        d7 = np.zeros((det.shape[0], 7), np.float64)
        d7[:, :4] = boxes
        d7[:, 4:6] = det[:, 4:6]
        d7[:, 6] = s
        per_scale.append(d7)
    nonempty = [d for d in per_scale if len(d)]
    if not nonempty:
        return np.array([]), per_scale  # msfd:285
    merged = np.vstack(nonempty)
    return weighted_nms(merged, len(heads_per_scale), iou_thres), per_scale


# ---- next row: widerface_evaluate/box_overlaps.pyx:15-55 ----------------------------------------
def bbox_overlaps(boxes, query_boxes):
    boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
    query_boxes = np.ascontiguousarray(query_boxes, np.float64).reshape(-1, 4)
    out = np.zeros((boxes.shape[0], query_boxes.shape[0]), np.float64)
    lib().fo_bbox_overlaps(_p(boxes), C.c_int64(boxes.shape[0]), _p(query_boxes), C.c_int64(query_boxes.shape[0]), _p(out))
    return out


# ---- "next" row: test-time augmentation around the backbone --------------------------------
def scale_img(img, ratio=1.0, same_shape=False, gs=32, flip=None):
    """utils/torch_utils.py:247-257 (+ the x.flip(fi) of models/yolo.py:368) on an NCHW float32 array: ATen's
    upsample_bilinear2d with align_corners=False and an explicit size (scale = in / out in fp32, source index
    clamped at 0), then F.pad(value=0.447) to the next multiple of gs."""
    import math
    img = np.asarray(img, np.float32)
    if flip:
        img = np.flip(img, axis=flip)
    if ratio == 1.0:
        return np.ascontiguousarray(img)
    n, c, h, w = img.shape
    rh, rw = int(h * ratio), int(w * ratio)
    oh, ow = (h, w) if same_shape else [math.ceil(v * ratio / gs) * gs for v in (h, w)]

    def taps(n_in, n_out):
        s = np.float32(n_in) / np.float32(n_out)
        d = (np.arange(n_out, dtype=np.float32) + np.float32(0.5)).astype(np.float64)
        f = np.maximum((np.float64(s) * d - 0.5).astype(np.float32), np.float32(0))  # fmaf(s, d + 0.5, -0.5): ATen's build contracts it
        i0 = f.astype(np.int32)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (f - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, (np.float32(1) - l1).astype(np.float32), l1
    y0, y1, hy, ly = taps(h, rh)
    x0, x1, hx, lx = taps(w, rw)
    top = hx * img[:, :, y0][:, :, :, x0] + lx * img[:, :, y0][:, :, :, x1]
    bot = hx * img[:, :, y1][:, :, :, x0] + lx * img[:, :, y1][:, :, :, x1]
    res = hy[:, None] * top + ly[:, None] * bot
    out = np.full((n, c, oh, ow), np.float32(0.447), np.float32)
    out[:, :, :rh, :rw] = res
    return out


def descale_pred(p, flips, scale, img_size):
    """models/yolo.py:402-417 on a (bs, rows, no) float32 array (returns a new array)."""
    p = np.array(p, np.float32, copy=True)
    p[..., :4] = p[..., :4] / np.float32(scale)
    if flips == 2:
        p[..., 1] = np.float32(img_size[0]) - p[..., 1]
    elif flips == 3:
        p[..., 0] = np.float32(img_size[1]) - p[..., 0]
    return p
