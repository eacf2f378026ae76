"""torch-tensor front end of libfdms: thin wrappers that validate shapes, allocate outputs with torch
(device memory + streams are PyTorch's job here, nothing else is) and pass raw pointers to the C ABI.
Every function enqueues on torch's current stream and returns without synchronising."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FDMS_CHW, FDMS_F16, FDMS_F32, FDMS_HWC, FDMS_U8, IOU_NCNN, IOU_TV_CPU, IOU_TV_CUDA, MAX_NMS  # noqa: F401

NO = 21

_DT = {torch.uint8: FDMS_U8, torch.float16: FDMS_F16, torch.float32: FDMS_F32}


def _dev(t):
    if not t.is_cuda:
        raise ValueError("libfdms operates on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ctx(dev):
    """The calling thread's native context for torch's current stream on `dev` (one context serves one stream)."""
    return _lib.ctx(dev, torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# ---------------------------------------------------------------------------------------------- pyramid
def pyramid(src, geoms, out_dtype=torch.float16, layout="chw", flip_rb=True, border=114, virt_hw=None, outs=None,
            ctx=None):
    """Fused letterbox pyramid.  src: (B, H, W, 3) uint8 CUDA (any row pitch: stride(1) bytes).
    geoms: list of _lib.Geom (one per scale, from letterbox_geometry on the virtual size).
    Returns one tensor per scale: (B, 3, h, w) for layout 'chw' or (B, h, w, 3) for 'hwc'."""
    if src.dim() != 4 or src.shape[3] != 3 or src.dtype != torch.uint8:
        raise ValueError("pyramid: src must be (B, H, W, 3) uint8")
    if src.stride(3) != 1 or src.stride(2) != 3:
        raise ValueError("pyramid: pixels must be contiguous (HWC, 3 bytes per pixel)")
    dev = _dev(src)
    B, H, W, _ = src.shape
    vh, vw = virt_hw if virt_hw is not None else (H, W)
    n = len(geoms)
    arr = (_lib.ScaleOut * n)()
    result = []
    for s, g in enumerate(geoms):
        shape = (B, 3, g.dst_h, g.dst_w) if layout == "chw" else (B, g.dst_h, g.dst_w, 3)
        if outs is not None:
            o = outs[s]
            if tuple(o.shape) != shape or o.dtype != out_dtype or not o.is_contiguous() or o.device != src.device:
                raise ValueError("pyramid: bad preallocated output %d" % s)
        else:
            o = torch.empty(shape, dtype=out_dtype, device=src.device)
        result.append(o)
        arr[s].geom = g
        arr[s].out = o.data_ptr()
        arr[s].out_img_stride = 3 * g.dst_h * g.dst_w
        arr[s].out_dtype = _DT[out_dtype]
        arr[s].out_layout = FDMS_CHW if layout == "chw" else FDMS_HWC
    c = ctx if ctx is not None else _ctx(dev)
    rc = _lib.lib().fdms_pyramid(c.handle, _ptr(src), B, src.stride(0) if B > 1 else H * src.stride(1), src.stride(1),
                                 H, W, vh, vw, arr, n, int(bool(flip_rb)), int(border), _stream(dev))
    c.check(rc, "fdms_pyramid")
    return result


def debug_poison_smem(byte_value=0xFF, device=None):
    """Test hook: fill the shared memory of every SM with one byte value (see include/fdms.h)."""
    dev = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    c = _ctx(dev)
    c.check(_lib.lib().fdms_debug_poison_smem(c.handle, int(byte_value), _stream(dev)), "fdms_debug_poison_smem")


# ---------------------------------------------------------------------------------------------- JPEG ingest
def jpeg_probe(data):
    """Header probe of a JPEG byte string (host only): _lib.JpegDesc with width / height / `supported` / `reason`, or None
    when the bytes are not a JPEG stream at all."""
    d = _lib.JpegDesc()
    rc = _lib.lib().fdms_jpeg_probe(data, len(data), C.byref(d))
    return d if rc == 0 else None


def jpeg_decode_batch(datas, hw, out=None, threads=0, device=None):
    """Device JPEG ingest: `datas` = list of JPEG byte strings of ONE image size `hw` -> ((n, h, w, 3) uint8 BGR CUDA tensor,
    status int32 array).  The Huffman decode runs on `threads` host threads (0 = all) inside the call; IDCT, chroma
    upsampling and colour conversion are enqueued on torch's current stream.  status[i] == 0: slot i holds cv2.imread's
    pixels; 1: the stream needs the host decoder (progressive, 4:1:1, EXIF rotation ...) and slot i is untouched;
    < 0: corrupt / wrong size."""
    n = len(datas)
    h, w = int(hw[0]), int(hw[1])
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if out is None:
        out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    elif tuple(out.shape) != (n, h, w, 3) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise ValueError("jpeg_decode_batch: `out` must be a contiguous (n, h, w, 3) uint8 CUDA tensor")
    di = _dev(out)
    ptrs = (C.c_char_p * n)(*datas)
    lens = (C.c_size_t * n)(*[len(b) for b in datas])
    status = np.zeros((n,), np.int32)
    c = _ctx(di)
    rc = _lib.lib().fdms_jpeg_decode_batch(c.handle, ptrs, lens, n, int(threads), _ptr(out), out.stride(0), out.stride(1), h, w,
                                           C.c_void_p(status.ctypes.data), _stream(di))
    c.check(rc, "fdms_jpeg_decode_batch")
    return out, status


# ---------------------------------------------------------------------------------------------- decode
def _level_table(scales_levels, strides, anchors_px):
    """scales_levels: list over scales of list over levels of (bs, 63, ny, nx) tensors."""
    anchors_px = np.asarray(anchors_px, np.float32).reshape(-1, 6)
    n = sum(len(lv) for lv in scales_levels)
    if n > _lib.MAX_LEVELS:
        raise ValueError("too many levels")
    arr = (_lib.Level * n)()
    rows_per_scale = []
    i = 0
    bs = scales_levels[0][0].shape[0]
    hold = []
    for s, levels in enumerate(scales_levels):
        off = 0
        for li, t in enumerate(levels):
            if t.dim() != 4 or t.shape[1] != 3 * NO or t.shape[0] != bs or t.dtype not in (torch.float16, torch.float32):
                raise ValueError("head level must be (bs, 63, ny, nx) fp16/fp32, got %r %s" % (tuple(t.shape), t.dtype))
            if not t.is_contiguous():
                t = t.contiguous()
            hold.append(t)
            arr[i].data = t.data_ptr()
            arr[i].dtype = _DT[t.dtype]
            arr[i].ny, arr[i].nx = t.shape[2], t.shape[3]
            arr[i].scale = s
            arr[i].row_offset = off
            arr[i].stride = float(strides[li])
            for k in range(6):
                arr[i].anchors[k] = float(anchors_px[li, k])
            off += 3 * t.shape[2] * t.shape[3]
            i += 1
        rows_per_scale.append(off)
    return arr, n, bs, rows_per_scale, hold


def decode_dense(levels, strides, anchors_px):
    """models/yolo.py:267-308: list of raw (bs, 63, ny, nx) -> pred (bs, rows, 21) fp32."""
    dev = _dev(levels[0])
    arr, n, bs, rows, hold = _level_table([levels], strides, anchors_px)
    pred = torch.empty((bs, rows[0], NO), dtype=torch.float32, device=levels[0].device)
    c = _ctx(dev)
    c.check(_lib.lib().fdms_decode_dense(c.handle, arr, n, bs, _ptr(pred), rows[0], _stream(dev)), "fdms_decode_dense")
    return pred


def decode_dense_head(levels, strides, anchors_px, na, no, nkpt):
    """The inference branch of ANY reference head (models/yolo.py Detect :59-106, IDetect :152-207, IKeypoint :267-308):
    list of raw (bs, na*no, ny, nx) fp16/fp32 CUDA tensors -> pred (bs, rows, no) fp32.  nkpt > 0: keypoint head
    (no == 6 + 3*nkpt, nc = 1); nkpt == 0: plain head (no == 5 + nc).  anchors_px: (nl, na, 2) pixels."""
    na, no, nkpt = int(na), int(no), int(nkpt)
    n = len(levels)
    if n == 0 or n > _lib.MAX_LEVELS:
        raise ValueError("decode_dense_head: 1..%d levels" % _lib.MAX_LEVELS)
    if not (1 <= na <= _lib.MAX_ANCHORS) or not (6 <= no <= _lib.MAX_HEAD_NO):
        raise NotImplementedError("decode_dense_head: na in 1..%d and no in 6..%d (got na=%d no=%d)"
                                  % (_lib.MAX_ANCHORS, _lib.MAX_HEAD_NO, na, no))
    if nkpt < 0 or (nkpt > 0 and no != 6 + 3 * nkpt):
        raise ValueError("decode_dense_head: a keypoint head has no = 6 + 3*nkpt columns (got no=%d nkpt=%d)" % (no, nkpt))
    anchors = np.ascontiguousarray(np.asarray(anchors_px, np.float32).reshape(n, na, 2))
    bs = levels[0].shape[0]
    arr = (_lib.Level * n)()
    hold, off = [], 0
    for i, t in enumerate(levels):
        if t.dim() != 4 or t.shape[1] != na * no or t.shape[0] != bs or t.dtype not in (torch.float16, torch.float32):
            raise ValueError("head level must be (bs, %d, ny, nx) fp16/fp32, got %r %s" % (na * no, tuple(t.shape), t.dtype))
        if not t.is_cuda:
            raise ValueError("decode_dense_head: CUDA tensors only (libfdms has no CPU path)")
        if not t.is_contiguous():
            t = t.contiguous()
        hold.append(t)
        arr[i].data = t.data_ptr()
        arr[i].dtype = _DT[t.dtype]
        arr[i].ny, arr[i].nx = t.shape[2], t.shape[3]
        arr[i].scale = 0
        arr[i].row_offset = off
        arr[i].stride = float(strides[i])
        off += na * t.shape[2] * t.shape[3]
    dev = _dev(levels[0])
    pred = torch.empty((bs, off, no), dtype=torch.float32, device=levels[0].device)
    c = _ctx(dev)
    c.check(_lib.lib().fdms_decode_dense_head(c.handle, arr, n, bs, na, no, nkpt, C.c_void_p(anchors.ctypes.data),
                                              _ptr(pred), off, _stream(dev)), "fdms_decode_dense_head")
    return pred


DECODE_BOXES_ONLY = 1  # include/fdms.h FDMS_DECODE_BOXES_ONLY
POST_SMALL_ONLY = 1    # include/fdms.h FDMS_POST_SMALL_ONLY
SMALL_LIMIT = 1024     # include/fdms.h FDMS_SMALL_LIMIT


class Candidates:
    """Device-side candidate lists of n_seg segments (see include/fdms.h fdms_decode_filter)."""

    def __init__(self, n_seg, cap, device, no=NO):
        self.n_seg, self.cap = n_seg, cap
        self.cand = torch.empty((n_seg, cap, no), dtype=torch.float32, device=device)
        self.row = torch.empty((n_seg, cap), dtype=torch.int32, device=device)
        self.count = torch.zeros((n_seg,), dtype=torch.int32, device=device)
        self.keep = torch.empty((n_seg, cap), dtype=torch.int32, device=device)
        self.n_keep = torch.zeros((n_seg,), dtype=torch.int32, device=device)


def decode_filter(scales_levels, strides, anchors_px, conf_thres, cap, out=None, landmarks=True):
    """Fused decode + filter + compaction over all scales of a batch.  Segment = b * n_scales + s.
    landmarks=False leaves the 15 landmark columns of the candidate rows unread and unwritten (what the multi-scale
    path needs: multi_scale_face_detector.py:151-153 drops them before the merge)."""
    dev = _dev(scales_levels[0][0])
    arr, n, bs, _, hold = _level_table(scales_levels, strides, anchors_px)
    ns = len(scales_levels)
    if out is None:
        out = Candidates(bs * ns, cap, scales_levels[0][0].device)
    c = _ctx(dev)
    rc = _lib.lib().fdms_decode_filter_ex(c.handle, arr, n, ns, bs, float(np.float32(conf_thres)), _ptr(out.cand),
                                          _ptr(out.row), _ptr(out.count), out.cap, 0 if landmarks else DECODE_BOXES_ONLY,
                                          _stream(dev))
    c.check(rc, "fdms_decode_filter_ex")
    return out


def filter_pred(pred, conf_thres, cap, out=None):
    """Confidence filter + xywh2xyxy on a decoded pred (bs, rows, 6 + 3*kpt) fp16/fp32 (nc == 1)."""
    dev = _dev(pred)
    if pred.dim() != 3 or pred.shape[2] < 6 or (pred.shape[2] - 6) % 3 or pred.dtype not in (torch.float16, torch.float32):
        raise ValueError("filter_pred: pred must be (bs, rows, 6 + 3*kpt) fp16/fp32")
    pred = pred.contiguous()
    bs, rows, no = pred.shape
    if out is None:
        out = Candidates(bs, cap, pred.device, no)
    c = _ctx(dev)
    rc = _lib.lib().fdms_filter_pred(c.handle, _ptr(pred), _DT[pred.dtype], bs, rows, no, float(np.float32(conf_thres)),
                                     _ptr(out.cand), _ptr(out.row), _ptr(out.count), out.cap, _stream(dev))
    c.check(rc, "fdms_filter_pred")
    return out


# ---------------------------------------------------------------------------------------------- nms
def nms_segments(rows, count, cap, iou_thres, iou_mode=IOU_TV_CUDA, max_nms=MAX_NMS, tie=None, box_off=0, score_off=4,
                 keep=None, n_keep=None):
    """rows: (n_seg, cap, row_stride) fp32; count (n_seg,) int32.  Returns (keep (n_seg, cap) int32, n_keep)."""
    dev = _dev(rows)
    n_seg, cap_, stride = rows.shape
    assert cap_ == cap and rows.dtype == torch.float32 and rows.is_contiguous()
    if keep is None:
        keep = torch.empty((n_seg, cap), dtype=torch.int32, device=rows.device)
    if n_keep is None:
        n_keep = torch.zeros((n_seg,), dtype=torch.int32, device=rows.device)
    c = _ctx(dev)
    rc = _lib.lib().fdms_nms(c.handle, _ptr(rows), stride, box_off, score_off, _ptr(tie), _ptr(count), n_seg, cap,
                             float(iou_thres), int(iou_mode), int(max_nms), _ptr(keep), _ptr(n_keep), _stream(dev))
    c.check(rc, "fdms_nms")
    return keep, n_keep


def nms(boxes, scores, iou_thres, iou_mode=IOU_TV_CUDA):
    """Single-segment torchvision.ops.nms replacement: returns int64 keep indices in score order."""
    dev = _dev(boxes)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    rows = torch.empty((1, n, 5), dtype=torch.float32, device=boxes.device)
    rows[0, :, :4] = boxes.float()
    rows[0, :, 4] = scores.float()
    count = torch.full((1,), n, dtype=torch.int32, device=boxes.device)
    keep, n_keep = nms_segments(rows, count, n, iou_thres, iou_mode, max_nms=0)
    nk = int(n_keep.item())
    if nk < 0:
        raise _lib.FdmsError("fdms_nms: segment exceeds the suppression-mask budget")
    return keep[0, :nk].long()


def gather_rows(rows, keep, n_keep, n_cols, out_cap):
    dev = _dev(rows)
    n_seg, cap, stride = rows.shape
    out = torch.empty((n_seg, out_cap, n_cols), dtype=torch.float32, device=rows.device)
    c = _ctx(dev)
    rc = _lib.lib().fdms_gather_rows(c.handle, _ptr(rows), stride, n_cols, _ptr(keep), _ptr(n_keep), n_seg, cap,
                                     _ptr(out), out_cap, _stream(dev))
    c.check(rc, "fdms_gather_rows")
    return out


# ---------------------------------------------------------------------------------------------- rescale
def _rows_view(coords, min_cols):
    if coords.dim() != 2 or coords.dtype != torch.float32 or coords.shape[1] < min_cols or coords.stride(1) != 1:
        raise ValueError("coords must be a 2-D fp32 tensor with unit column stride")
    return coords.shape[0], coords.shape[1], coords.stride(0) if coords.shape[0] > 1 else coords.shape[1]


def scale_coords_(img1_shape, coords, img0_shape, kpt_step=0, ratio_pad=None):
    """In place utils/general.py:374-404 on a CUDA fp32 (n, cols) tensor/view.  ratio_pad = (gain, (pad_x, pad_y))
    replaces the gain / padding derived from the two shapes (:379-383)."""
    n, cols, rs = _rows_view(coords, 4 if kpt_step == 0 else 2)
    if n == 0:
        return coords
    dev = _dev(coords)
    c = _ctx(dev)
    if ratio_pad is None:
        rc = _lib.lib().fdms_scale_coords(c.handle, _ptr(coords), n, cols, rs, int(img1_shape[0]), int(img1_shape[1]),
                                          int(img0_shape[0]), int(img0_shape[1]), int(kpt_step), _stream(dev))
    else:
        gain, pad = ratio_pad
        rc = _lib.lib().fdms_scale_coords_ratio_pad(c.handle, _ptr(coords), n, cols, rs, float(gain), float(pad[0]),
                                                    float(pad[1]), int(img0_shape[0]), int(img0_shape[1]),
                                                    int(kpt_step), _stream(dev))
    c.check(rc, "fdms_scale_coords")
    return coords


def scale_coords_api_(img_input_shape, boxes, img0_shape, do_round=False):
    """In place utils/preprocess_yolo_predict.py:122-157 on a CUDA fp32 (n, >=4) tensor/view."""
    n, cols, rs = _rows_view(boxes, 4)
    if n == 0:
        return boxes
    dev = _dev(boxes)
    c = _ctx(dev)
    rc = _lib.lib().fdms_scale_coords_api(c.handle, _ptr(boxes), n, rs, int(img_input_shape[0]), int(img0_shape[0]),
                                          int(img0_shape[1]), int(bool(do_round)), _stream(dev))
    c.check(rc, "fdms_scale_coords_api")
    return boxes


# ---------------------------------------------------------------------------------------------- merge / post
class Detections:
    """(n_img, out_cap, 7) fp32 dets [x1,y1,x2,y2,conf,cls,scale_idx] + n_out, device resident."""

    def __init__(self, n_img, out_cap, device, with_kpts=False):
        self.n_img, self.out_cap = n_img, out_cap
        self.dets = torch.empty((n_img, out_cap, 7), dtype=torch.float32, device=device)
        self.n_out = torch.zeros((n_img,), dtype=torch.int32, device=device)
        self.merged_in = torch.zeros((n_img,), dtype=torch.int32, device=device)
        self.kpts = torch.empty((n_img, out_cap, 15), dtype=torch.float32, device=device) if with_kpts else None


def merge_scales(cands, n_img, n_scales, in_h, orig_hw, iou_thres, out_cap, with_kpts=False, out=None):
    dev = _dev(cands.cand)
    if out is None:
        out = Detections(n_img, out_cap, cands.cand.device, with_kpts)
    in_h_arr = (C.c_int32 * n_scales)(*[int(h) for h in in_h])
    c = _ctx(dev)
    rc = _lib.lib().fdms_merge_scales(c.handle, _ptr(cands.cand), _ptr(cands.keep), _ptr(cands.n_keep), n_img, n_scales,
                                      cands.cap, in_h_arr, int(orig_hw[0]), int(orig_hw[1]), float(iou_thres),
                                      _ptr(out.dets), _ptr(out.n_out), out.out_cap, _ptr(out.merged_in), _ptr(out.kpts),
                                      _stream(dev))
    c.check(rc, "fdms_merge_scales")
    return out


def weighted_rows(dets64, n_scales, want_weights=False):
    """(n,7) float64 CUDA detections -> (1, n, 8) fp32 merge rows (weighted score in column 4)
    [and the (n,) float64 weights of calculate_scale_weights]."""
    dev = _dev(dets64)
    n = dets64.shape[0]
    assert dets64.dtype == torch.float64 and dets64.is_contiguous() and dets64.shape[1] == 7
    rows = torch.empty((1, max(n, 1), 8), dtype=torch.float32, device=dets64.device)
    c = _ctx(dev)
    w = torch.empty((n,), dtype=torch.float64, device=dets64.device) if want_weights else None
    c.check(_lib.lib().fdms_weighted_rows(c.handle, _ptr(dets64), n, int(n_scales), _ptr(rows), _ptr(w), _stream(dev)),
            "fdms_weighted_rows")
    return (rows, w) if want_weights else rows


class PostPlan:
    """A frozen post-processing step (decode_filter -> NMS -> merge) over fixed buffers: one C call per
    run, so the launches go out back to back (and can be captured in a CUDA graph)."""

    def __init__(self, scales_levels, strides, anchors_px, in_h, orig_hw, conf_thres, iou_thres, cap, out_cap,
                 iou_mode=IOU_TV_CUDA, with_kpts=False, ctx=None, small_only=False):
        """small_only: launch only the single-CTA kernels (FDMS_POST_SMALL_ONLY): correct iff afterwards every
        `cands.count` and `out.merged_in` is <= SMALL_LIMIT, which the caller must check (`small_ok`) -- otherwise run
        the same plan again with `small_only = False`."""
        self.dev = _dev(scales_levels[0][0])
        device = scales_levels[0][0].device
        self._strides, self._anchors = list(strides), np.asarray(anchors_px, np.float32)
        self.arr, self.n_levels, self.bs, self.rows_per_scale, self._hold = _level_table(scales_levels, strides,
                                                                                        anchors_px)
        self.n_scales = len(scales_levels)
        self.cands = Candidates(self.bs * self.n_scales, cap, device)
        self.out = Detections(self.bs, out_cap, device, with_kpts)
        self._in_h = (C.c_int32 * self.n_scales)(*[int(h) for h in in_h])
        d = _lib.PostDesc()
        d.levels = C.cast(self.arr, C.POINTER(_lib.Level))
        d.n_levels, d.n_scales, d.bs, d.cap = self.n_levels, self.n_scales, self.bs, cap
        d.conf_thres = float(np.float32(conf_thres))
        d.iou_mode = int(iou_mode)
        d.iou_thres = float(iou_thres)
        d.in_h_host = C.cast(self._in_h, C.POINTER(C.c_int32))
        # orig_hw: (H0, W0) for the whole batch, or -- ragged batch -- a (bs, 2) int32 CUDA tensor with one pair per image
        if isinstance(orig_hw, torch.Tensor):
            if orig_hw.dtype != torch.int32 or tuple(orig_hw.shape) != (self.bs, 2) or not orig_hw.is_cuda or not orig_hw.is_contiguous():
                raise ValueError("PostPlan: per-image orig_hw must be a contiguous (bs, 2) int32 CUDA tensor")
            self._orig_hw = orig_hw
            d.orig_h, d.orig_w, d.orig_hw_dev = 1, 1, orig_hw.data_ptr()
        else:
            self._orig_hw = None
            d.orig_h, d.orig_w, d.orig_hw_dev = int(orig_hw[0]), int(orig_hw[1]), None
        d.cand, d.cand_row, d.count = self.cands.cand.data_ptr(), self.cands.row.data_ptr(), self.cands.count.data_ptr()
        d.keep, d.n_keep = self.cands.keep.data_ptr(), self.cands.n_keep.data_ptr()
        d.dets, d.n_out, d.out_cap = self.out.dets.data_ptr(), self.out.n_out.data_ptr(), out_cap
        d.merged_in = self.out.merged_in.data_ptr()
        d.kpts_out = self.out.kpts.data_ptr() if with_kpts else None
        d.flags = POST_SMALL_ONLY if small_only else 0
        self.desc = d
        self.ctx = ctx if ctx is not None else _ctx(self.dev)

    def rebind(self, scales_levels):
        """Point the plan at a new set of raw head tensors of the same shapes (buffers are reused)."""
        arr, n, bs, rows, hold = _level_table(scales_levels, self._strides,
                                              self._anchors)
        assert n == self.n_levels and bs == self.bs
        self.arr, self._hold = arr, hold
        self.desc.levels = C.cast(self.arr, C.POINTER(_lib.Level))

    @property
    def small_only(self):
        return bool(self.desc.flags & POST_SMALL_ONLY)

    @small_only.setter
    def small_only(self, on):
        self.desc.flags = (self.desc.flags | POST_SMALL_ONLY) if on else (self.desc.flags & ~POST_SMALL_ONLY)

    @staticmethod
    def small_ok(max_count, max_merged_in):
        """The check a small_only run needs: nothing exceeded the single-CTA kernels' limit."""
        return int(max_count) <= SMALL_LIMIT and int(max_merged_in) <= SMALL_LIMIT

    def run(self):
        rc = _lib.lib().fdms_postprocess(self.ctx.handle, C.byref(self.desc), _stream(self.dev))
        self.ctx.check(rc, "fdms_postprocess")
        return self.out

    @property
    def launches(self):
        """Kernel launches of one run() (memset excluded): decode filter (+ candidate decode), the per-scale NMS (one
        fused launch when no segment can exceed 1024 boxes, else + sort + greedy sweep), the cross-scale merge (one
        fused launch when no image can have more than 1024 merged rows, else + sort + greedy sweep + emit)."""
        import os
        small = SMALL_LIMIT
        dec = 2 if os.environ.get("FDMS_DECODE_FUSED") == "0" else 1  # the filter CTAs decode their own survivors
        if self.small_only:
            return dec + 2
        mcap = min(self.n_scales * self.cands.cap, self.out.out_cap)
        return dec + (1 if self.cands.cap <= small else 3) + (1 if mcap <= small else 4)
