"""Drop-in for the reference's `utils/datasets.py::letterbox` (utils/datasets.py:873-903) on libfdms.

Same signature, same return value `(img, ratio, (dw, dh))`, numpy HWC uint8 in and out; the resize
(cv2.INTER_LINEAR fixed point, bit exact) and the constant border run in the fused pyramid kernel.
"""
import numpy as np
import torch

from . import _lib, ops


def _to_device_u8(img, device):
    if isinstance(img, torch.Tensor):
        t = img
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError("letterbox: expected an (H, W, 3) uint8 image")
        return t.to(device, non_blocking=True).contiguous()
    a = np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("letterbox: expected an (H, W, 3) uint8 image, got %r %s" % (a.shape, a.dtype))
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


def letterbox(img, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32,
              device=None):
    """utils/datasets.py:873.  Returns (ndarray HWC uint8, (rw, rh), (dw, dh)) like the reference.
    A torch CUDA uint8 image is accepted too and then a CUDA tensor is returned (no host round trip)."""
    if len(set(int(c) for c in color)) != 1:
        raise NotImplementedError("letterbox: only a grey border colour (r == g == b) is supported")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    on_device = isinstance(img, torch.Tensor) and img.is_cuda
    t = _to_device_u8(img, img.device if on_device else dev)
    h, w = t.shape[:2]
    g = _lib.letterbox_geometry(h, w, new_shape, auto=auto, scaleFill=scaleFill, scaleup=scaleup, stride=stride)
    (out,) = ops.pyramid(t[None], [g], out_dtype=torch.uint8, layout="hwc", flip_rb=False, border=int(color[0]))
    out = out[0]
    ratio, pad = (g.ratio_w, g.ratio_h), (g.dw, g.dh)
    if on_device:
        return out, ratio, pad
    return out.cpu().numpy(), ratio, pad


def letterbox_shape(src_hw, new_shape=640, auto=True, scaleup=True, stride=32):
    """(H, W) of the letterboxed image without touching pixels."""
    g = _lib.letterbox_geometry(src_hw[0], src_hw[1], new_shape, auto=auto, scaleup=scaleup, stride=stride)
    return g.dst_h, g.dst_w
