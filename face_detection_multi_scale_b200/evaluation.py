"""Drop-in for `widerface_evaluate/box_overlaps.pyx::bbox_overlaps` (:15-55) on libfdms.
Same signature: (N,4) and (K,4) float64 ndarrays -> (N,K) float64 IoU matrix, +1 pixel convention."""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops


def bbox_overlaps(boxes, query_boxes, device=None):
    on_dev = isinstance(boxes, torch.Tensor) and boxes.is_cuda
    dev = boxes.device if on_dev else torch.device("cuda", torch.cuda.current_device() if device is None else device)
    b = boxes if on_dev else torch.from_numpy(np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)).to(dev)
    q = query_boxes if isinstance(query_boxes, torch.Tensor) and query_boxes.is_cuda else \
        torch.from_numpy(np.ascontiguousarray(query_boxes, np.float64).reshape(-1, 4)).to(dev)
    if b.dtype != torch.float64 or q.dtype != torch.float64:
        raise ValueError("bbox_overlaps: float64 boxes expected (DTYPE = np.float in the reference)")
    b, q = b.contiguous(), q.contiguous()
    out = torch.zeros((b.shape[0], q.shape[0]), dtype=torch.float64, device=dev)
    c = ops._ctx(dev.index)
    rc = _lib.lib().fdms_bbox_overlaps(c.handle, C.c_void_p(b.data_ptr()), b.shape[0], C.c_void_p(q.data_ptr()), q.shape[0],
                                       C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    c.check(rc, "fdms_bbox_overlaps")
    return out if on_dev else out.cpu().numpy()
