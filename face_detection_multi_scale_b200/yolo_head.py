"""The head decode of `models/yolo.py` (IKeypoint.forward :253-308; same arithmetic as Detect :59-106 and
IDetect :152-207) on libfdms.  The 1x1 convolutions of the head stay in PyTorch (north_star: the backbone
forward is unchanged); everything after them -- view/permute/contiguous/clone, sigmoid, grid + anchor decode,
landmark decode, cat -- is one kernel reading the conv outputs in place.

    head_convs(head, x)        the conv part of IKeypoint/IDetect/Detect.forward (yolo.py:267-271), NCHW out
    decode_head(...)           raw (bs, 63, ny, nx) list -> pred (bs, rows, 21)   [compat level]
    patch_head(head)           replace head.forward: returns (pred, x) like the reference in eval mode and
                               keeps the raw NCHW maps in head.fdms_raw for the fused path
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
        if getattr(head, "nkpt", None):
            det = torch.cat((det, head.m_kpt[i](xi)), 1)
        out.append(det)
    return out


def decode_head(raw_levels, strides, anchors_px):
    """list of raw (bs, 63, ny, nx) CUDA tensors -> pred (bs, rows, 21) fp32, rows level-major then anchor, y, x
    (models/yolo.py:306-308)."""
    return ops.decode_dense(raw_levels, strides, anchors_px)


def _forward(self, x):
    if self.training:
        raise RuntimeError("the libfdms head replaces the inference branch only (call model.eval())")
    raw = head_convs(self, x)
    self.fdms_raw = raw
    if getattr(self._fdms_tls, "fused", False):  # fused level (per calling thread): the detector consumes the raw maps
        return raw
    strides, anchors = self._fdms_geom
    pred = decode_head(raw, strides, anchors)
    if raw[0].dtype != torch.float32:
        pred = pred.to(raw[0].dtype)
    xs = []
    for r in raw:  # second element of the reference's return value: (bs, na, ny, nx, no) raw maps
        bs, _, ny, nx = r.shape
        xs.append(r.view(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2))
    return pred, xs


def patch_head(head):
    """Route `head.forward` through libfdms.  Returns the head.  Requires no=21 (nc=1, nkpt=5), na=3."""
    if head.no != ops.NO or head.na != 3:
        raise NotImplementedError("libfdms decodes the face head only (na=3, no=21); got na=%d no=%d" % (head.na, head.no))
    head._fdms_geom = head_geometry(head)
    head.fdms_raw = None
    head._fdms_tls = threading.local()  # the fused/compat switch is per thread: one detector may serve a thread pool
    head.forward = types.MethodType(_forward, head)
    return head


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
