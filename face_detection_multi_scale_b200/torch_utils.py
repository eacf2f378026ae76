"""utils/torch_utils.py::scale_img (:247-257) and the test-time augmentation of models/yolo.py (Model.forward_augment
:363-374, Model._descale_pred :402-417) on libfdms: the elementwise work on either side of the backbone.  The backbone
forward itself (`forward_once`) stays the caller's PyTorch code.

    scale_img(img, ratio, same_shape=False, gs=32, flip=None)   drop-in (+ the flip forward_augment applies first)
    descale_pred(p, flips, scale, img_size, out=None, row_offset=0)
    forward_augment(forward_once, x, gs)                         -> (cat of the 3 de-scaled predictions, None)
"""
import math

import torch

from . import _lib, ops

TTA_SCALES = (1, 0.83, 0.67)   # models/yolo.py:365
TTA_FLIPS = (None, 3, None)    # models/yolo.py:366 (2 = up-down, 3 = left-right)


def scale_img(img, ratio=1.0, same_shape=False, gs=32, flip=None):
    """img (bs, c, h, w) CUDA f16/f32 scaled by `ratio` and padded with 0.447 to a multiple of `gs`; `flip` (2 / 3)
    reads the source flipped, i.e. scale_img(img.flip(flip), ...) in one pass."""
    if flip not in (None, 0, 2, 3):
        raise ValueError("flip must be None, 2 (up-down) or 3 (left-right)")
    if ratio == 1.0:
        return img.flip(flip) if flip else img
    dev = ops._dev(img)
    if img.dim() != 4 or img.dtype not in (torch.float16, torch.float32):
        raise ValueError("scale_img: img must be (bs, c, h, w) fp16/fp32")
    img = img.contiguous()
    bs, c, h, w = img.shape
    rs = (int(h * ratio), int(w * ratio))
    oh, ow = (h, w) if same_shape else [math.ceil(v * ratio / gs) * gs for v in (h, w)]  # same_shape: pad back to (h, w)
    if oh < rs[0] or ow < rs[1]:
        raise ValueError("scale_img: ratio > 1 with same_shape would crop (negative F.pad); not supported")
    out = torch.empty((bs, c, oh, ow), dtype=img.dtype, device=img.device)
    cx = ops._ctx(dev)
    rc = _lib.lib().fdms_scale_img(cx.handle, ops._ptr(img), _lib.FDMS_F16 if img.dtype == torch.float16 else _lib.FDMS_F32,
                                   bs, c, h, w, rs[0], rs[1], oh, ow, int(flip or 0), 0.447, ops._ptr(out), ops._stream(dev))
    cx.check(rc, "fdms_scale_img")
    return out


def descale_pred(p, flips, scale, img_size, out=None, row_offset=0):
    """models/yolo.py:402-417 on p (bs, rows, no) CUDA fp32.  With `out` (bs, rows_total, no) the result is written
    at `row_offset` (the concatenation of forward_augment); otherwise in place, like the reference's inplace branch."""
    dev = ops._dev(p)
    if p.dim() != 3 or p.dtype != torch.float32 or not p.is_contiguous():
        raise ValueError("descale_pred: p must be a contiguous (bs, rows, no) fp32 tensor")
    dst = p if out is None else out
    bs, rows, no = p.shape
    cx = ops._ctx(dev)
    rc = _lib.lib().fdms_descale_pred(cx.handle, ops._ptr(p), bs, rows, no, float(scale), int(flips or 0), int(img_size[0]),
                                      int(img_size[1]), ops._ptr(dst), dst.shape[1], int(row_offset), ops._stream(dev))
    cx.check(rc, "fdms_descale_pred")
    return dst


def forward_augment(forward_once, x, gs=32):
    """Model.forward_augment (models/yolo.py:363-374): `forward_once(xi)` must return the decoded prediction
    (bs, rows_i, no) first (a patched head does).  Returns (pred (bs, sum rows_i, no), None)."""
    img_size = x.shape[-2:]
    ys = []
    for si, fi in zip(TTA_SCALES, TTA_FLIPS):
        yi = forward_once(scale_img(x, si, gs=gs, flip=fi))
        yi = yi[0] if isinstance(yi, (tuple, list)) else yi
        ys.append(yi.float().contiguous())
    out = torch.empty((ys[0].shape[0], sum(y.shape[1] for y in ys), ys[0].shape[2]), dtype=torch.float32, device=x.device)
    r0 = 0
    for yi, si, fi in zip(ys, TTA_SCALES, TTA_FLIPS):
        descale_pred(yi, fi, si, img_size, out=out, row_offset=r0)
        r0 += yi.shape[1]
    return out, None
