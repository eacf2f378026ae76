"""The head decode of `models/yolo.py` (IKeypoint.forward :253-308; same arithmetic as Detect :59-106 and
IDetect :152-207) on libfdms.  The 1x1 convolutions of the head stay in PyTorch (north_star: the backbone
forward is unchanged); everything after them -- view/permute/contiguous/clone, sigmoid, grid + anchor decode,
landmark decode, cat -- is one kernel reading the conv outputs in place.

    head_convs(head, x)        the conv part of IKeypoint/IDetect/Detect.forward (yolo.py:267-271), NCHW out
    decode_head(...)           raw (bs, 63, ny, nx) list -> pred (bs, rows, 21)   [compat level]
    patch_head(head)           replace head.forward: returns (pred, x) like the reference in eval mode and
                               keeps the raw NCHW maps in head.fdms_raw for the fused path; any na / nc / nkpt
                               the reference's own forward can run (head_columns)
    head_geometry(head)        (strides, anchors in pixels) of a reference head module
"""
import threading
import types

import numpy as np
import torch

from . import ops


def head_geometry(head):
    """strides (nl,) and pixel anchors (nl, na, 2) of a reference Detect/IDetect/IKeypoint module.
    `anchor_grid` is in pixels; `stride` is a plain attribute."""
    strides = [float(s) for s in head.stride]
    anchors = head.anchor_grid.detach().float().reshape(head.nl, head.na, 2).cpu().numpy()
    return strides, anchors


def head_convs(head, x):
    """models/yolo.py:267-271: per level cat(im(m(ia(x))), m_kpt(x)) -> (bs, na*no, ny, nx).  Falls back to the
    plain Detect conv (m only) for heads without implicit layers."""
    out = []
    for i in range(head.nl):
        xi = x[i]
        if hasattr(head, "ia"):
            det = head.im[i](head.m[i](head.ia[i](xi)))
        else:
            det = head.m[i](xi)
        if getattr(head, "nkpt", None) and hasattr(head, "m_kpt"):  # (a head whose `m` delivers all na*no channels has none)
            det = torch.cat((det, head.m_kpt[i](xi)), 1)
        out.append(det)
    return out


def decode_head(raw_levels, strides, anchors_px, na=3, no=ops.NO, nkpt=5):
    """list of raw (bs, na*no, ny, nx) CUDA tensors -> pred (bs, rows, no) fp32, rows level-major then anchor, y, x
    (models/yolo.py:306-308).  The face head (na=3, no=21) takes the specialised kernel, every other head the
    reference can run in eval mode the general one (ops.decode_dense_head)."""
    if na == 3 and no == ops.NO and nkpt == 5:
        return ops.decode_dense(raw_levels, strides, anchors_px)
    return ops.decode_dense_head(raw_levels, strides, anchors_px, na, no, nkpt)


def head_columns(head):
    """(na, no, nkpt) of a reference head, refusing the combinations the reference itself cannot run in eval mode:
      * keypoints with nc != 1 -- every head slices `x[i][..., :6]` (models/yolo.py:67,160,279), the keypoint grids no
        longer line up and the reference raises;
      * keypoints with inplace=False -- `y[..., 6:]` of the 6-column `x_det.sigmoid()` is empty (:97,192,301), raises;
      * nkpt == 0, nc > 1 with inplace=True -- `cat((xy, wh, y[..., 4:], x_kpt))` (:84,186,297) has 2*no - 6 columns
        and `y.view(bs, -1, no)` scrambles or raises; such heads run with inplace=False (:95-99) only."""
    na, no = int(head.na), int(head.no)
    nkpt = int(getattr(head, "nkpt", 0) or 0)
    nc = int(getattr(head, "nc", no - 5 - 3 * nkpt))
    inplace = bool(getattr(head, "inplace", True))
    if no != 5 + nc + 3 * nkpt:
        raise ValueError("head: no=%d is not 5 + nc + 3*nkpt (nc=%d, nkpt=%d)" % (no, nc, nkpt))
    if nkpt > 0 and nc != 1:
        raise NotImplementedError("keypoint heads decode nc=1 only: the reference slices [..., :6] (models/yolo.py:279)")
    if nkpt > 0 and not inplace:
        raise NotImplementedError("keypoint heads with inplace=False fail in the reference too (models/yolo.py:301)")
    if nkpt == 0 and nc > 1 and inplace:
        raise NotImplementedError("nc > 1 without keypoints needs inplace=False: the reference's inplace branch emits "
                                  "2*no - 6 columns (models/yolo.py:297)")
    return na, no, nkpt


def _forward(self, x):
    if self.training:
        raise RuntimeError("the libfdms head replaces the inference branch only (call model.eval())")
    raw = head_convs(self, x)
    self.fdms_raw = raw
    if getattr(self._fdms_tls, "fused", False):  # fused level (per calling thread): the detector consumes the raw maps
        return raw
    strides, anchors = self._fdms_geom
    na, no, nkpt = self._fdms_cols
    # float32 whatever the model's precision: the reference's own pred is a float32 tensor for a half model too --
    # `self.grid[i]` is created float32 (models/yolo.py:310-313), so the xy columns and then `torch.cat` promote the row
    pred = decode_head(raw, strides, anchors, na, no, nkpt)
    xs, off = [], 0
    for r in raw:  # second element of the reference's return value: (bs, na, ny, nx, no) raw maps
        bs, _, ny, nx = r.shape
        xi = r.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2)
        if self._fdms_mutates_kpt:  # Detect / IDetect decode the keypoint columns IN the returned tensor (:76-81, :169-183)
            xi = xi.contiguous()
            xi[..., 6:] = pred[:, off:off + na * ny * nx].view(bs, na, ny, nx, no)[..., 6:].to(xi.dtype)
        xs.append(xi)
        off += na * ny * nx
    return pred, xs


def patch_head(head):
    """Route `head.forward` through libfdms.  Returns the head.  Any Detect / IDetect / IKeypoint the reference can run
    in eval mode (see head_columns); the FUSED level of the detector needs the face head (na=3, no=21)."""
    head._fdms_cols = head_columns(head)
    if head.na > ops._lib.MAX_ANCHORS or head.no > ops._lib.MAX_HEAD_NO:
        raise NotImplementedError("libfdms decodes heads with na <= %d and no <= %d; got na=%d no=%d"
                                  % (ops._lib.MAX_ANCHORS, ops._lib.MAX_HEAD_NO, head.na, head.no))
    head._fdms_geom = head_geometry(head)
    head._fdms_mutates_kpt = head._fdms_cols[2] > 0 and type(head).__name__ in ("Detect", "IDetect")
    head.fdms_raw = None
    head._fdms_tls = threading.local()  # the fused/compat switch is per thread: one detector may serve a thread pool
    head.forward = types.MethodType(_forward, head)
    return head


def is_face_head(head):
    """The head geometry the fused decode + filter kernel is written for (every cfg of the reference)."""
    return int(head.na) == 3 and int(head.no) == ops.NO and int(getattr(head, "nkpt", 0) or 0) == 5


def set_fused(head, on):
    """Route this THREAD's next forward of `head` to the raw (fused-level) output."""
    head._fdms_tls.fused = bool(on)


def find_head(model):
    """The detection head of a reference `Model` (its last layer, models/yolo.py:344)."""
    m = model.module if hasattr(model, "module") else model
    seq = m.model if hasattr(m, "model") else m
    head = seq[-1] if hasattr(seq, "__getitem__") else seq
    for attr in ("nl", "na", "no", "anchor_grid", "stride", "m"):
        if not hasattr(head, attr):
            raise TypeError("model's last layer does not look like a YOLO detection head (missing .%s)" % attr)
    return head


def anchors_as_array(anchors_px):
    return np.asarray(anchors_px, np.float32)
