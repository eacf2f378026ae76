"""Drop-ins for the hot-path functions of the reference's `utils/general.py`, running on libfdms:

    non_max_suppression   utils/general.py:506-600   (kpt_label=5, nc=1 -- every cfg/*.yaml of the reference)
    scale_coords          utils/general.py:374-399   (in place, returns coords)
    clip_coords           utils/general.py:401-404
    xywh2xyxy             utils/general.py:299-306
    check_img_size        utils/general.py:130-135
    make_divisible        utils/general.py:210-212

Same names, argument meaning and return types as the reference.  Tensors that live on the CPU are
moved to the current CUDA device and back (there is no CPU implementation in this package).
"""
import math

import torch

from . import _lib, ops

# torchvision.ops.nms flavour reproduced by non_max_suppression: the reference calls it
# on whatever device the prediction lives on -- the CUDA kernel in production.
IOU_MODE = ops.IOU_TV_CUDA
MAX_NMS = 30000         # utils/general.py:520
_DEFAULT_CAP = 4096     # first-try candidate capacity per image; grown on overflow


def make_divisible(x, divisor):  # utils/general.py:210
    return math.ceil(x / divisor) * divisor


def check_img_size(img_size, s=32):  # utils/general.py:130
    new_size = make_divisible(img_size, int(s))
    if new_size != img_size:
        print('WARNING: --img-size %g must be multiple of max stride %g, updating to %g' % (img_size, s, new_size))
    return new_size


def _cuda(t):
    return t if t.is_cuda else t.to(torch.device("cuda", torch.cuda.current_device()))


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), kpt_label=5, nc=None, iou_mode=None):
    """utils/general.py:506.  prediction (bs, rows, no) with nc == 1 (no = 21 for the face head, 6 without
    keypoints) -> list of (n, no) tensors [x1, y1, x2, y2, conf, cls, kpts...] per image, (0, 6) zeros for images
    without detections (:527).
    One filter launch + one batched NMS (one launch for segments of <= 1024 boxes) + one gather for the whole batch; a single
    device->host read of the per-image counts.
    fp16 predictions (the reference's CUDA default) are up-cast: the filter, xywh->xyxy and the IoU run in fp32 and the
    kept rows are rounded back to fp16, whereas the reference does all of it in half arithmetic -- candidate sets and
    boxes can differ from its fp16 result in the last fp16 bit (tests/ref_gpu_check.py, part C reports the agreement)."""
    if prediction.dim() != 3:
        raise ValueError("non_max_suppression: prediction must be (bs, rows, no)")
    no = prediction.shape[2]
    if no == 6 and kpt_label and nc is None:
        # multi_scale_face_detector.py:134 (`pred[..., :6]` with the default kpt_label=5): the reference computes
        # nc = 6 - 5 - 15 < 0, its `x[:, 5:5+nc] *= x[:, 4:5]` is a no-op on an empty slice, and the confidence it filters
        # and sorts by is the RAW class score (rows with obj <= thres are dropped first, :515/:530).  Same thing here:
        # obj = 1 makes conf = cls, rows the objectness test drops get cls = -inf.
        p2 = prediction.clone()
        p2[..., 5] = torch.where(prediction[..., 4] > conf_thres, prediction[..., 5], torch.full_like(prediction[..., 5], float("-inf")))
        p2[..., 4] = 1.0
        return non_max_suppression(p2, conf_thres, iou_thres, classes, agnostic, multi_label, labels, kpt_label=0, nc=1,
                                   iou_mode=iou_mode)
    if nc is None:
        nc = prediction.shape[2] - 5 if not kpt_label else prediction.shape[2] - 5 - kpt_label * 3
    if nc > 1 and not kpt_label and no == 5 + nc:
        return _nms_multiclass(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, nc, iou_mode)
    if nc != 1 or no < 6 or (no - 6) % 3:
        raise NotImplementedError("non_max_suppression on libfdms covers single-class heads with keypoints (nc=1, no = 6 + "
                                  "3*kpt: the face head no=21), the [..., :6] fallback of msfd:134 and multi-class heads "
                                  "without keypoints (kpt_label=0, no = 5 + nc); got no=%d kpt_label=%r nc=%r"
                                  % (no, kpt_label, nc))
    if labels and no != 6:
        # :535-541 builds the label rows with nc + 5 columns; with keypoint columns torch.cat raises in the reference too
        raise ValueError("non_max_suppression: apriori `labels` need a prediction without keypoint columns (no = 6)")
    # multi_label &= nc > 1 (:522) -> always False here; agnostic only changes the class offset, 0 for nc == 1
    src_device, src_dtype = prediction.device, prediction.dtype
    pred = _cuda(prediction)
    if pred.dtype not in (torch.float16, torch.float32):
        pred = pred.float()
    bs, rows, _ = pred.shape
    empty = torch.zeros((0, 6), device=src_device)
    if bs == 0:
        return []
    if classes is not None and 0 not in [int(c) for c in (classes if hasattr(classes, "__iter__") else [classes])]:
        return [empty] * bs  # every detection has class 0 (:566-567)
    mode = IOU_MODE if iou_mode is None else iou_mode
    cap = min(rows, _DEFAULT_CAP)
    if labels:  # autolabelling (:533-541): room for the apriori boxes after the filtered predictions
        n_lab = [len(labels[b]) if b < len(labels) else 0 for b in range(bs)]
        cap = min(rows, _DEFAULT_CAP) + max(n_lab)
    while True:
        c = ops.filter_pred(pred, conf_thres, cap)
        if labels and max(n_lab):
            cap_needed = _append_labels(c, labels, n_lab, rows, conf_thres)
            if cap_needed > cap:
                cap = cap_needed
                continue
        ops.nms_segments(c.cand, c.count, cap, iou_thres, mode, max_nms=MAX_NMS, tie=c.row, keep=c.keep, n_keep=c.n_keep)
        counts = torch.stack((c.count, c.n_keep)).cpu()  # the one synchronisation of this call
        n_cand, n_keep = counts[0], counts[1]
        if int(n_cand.max()) > cap:  # more candidates than capacity: redo with room for all of them
            cap = min(rows, int(n_cand.max()))
            continue
        break
    if int(n_keep.min()) < 0:
        raise _lib.FdmsError("non_max_suppression: a segment exceeded the suppression-mask budget "
                             "(raise FDMS_MASK_BUDGET_MB)")
    kmax = int(n_keep.max())
    if kmax == 0:
        return [empty] * bs
    out = ops.gather_rows(c.cand, c.keep, c.n_keep, no, kmax)
    if out.dtype != src_dtype:
        out = out.to(src_dtype)
    if out.device != src_device:
        out = out.to(src_device)
    return [out[b, :int(n_keep[b])] if int(n_keep[b]) else empty for b in range(bs)]


def _nms_multiclass(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, nc, iou_mode):
    """utils/general.py:506-600 for nc > 1 without keypoints (the Detect / IDetect heads of the YOLO family the
    reference's code base derives from; none of its own cfg/*.yaml is multi-class): conf = obj * cls per class (:547),
    `multi_label` = one row per (box, class) above the threshold in nonzero() order (:553-555), else the best class
    (:558-559); `classes` filter (:566-567); boxes offset by class * 4096 unless `agnostic` (:582); one batched NMS launch
    for all images (torchvision semantics, `iou_mode`).  The filtering is torch tensor arithmetic in the prediction's own
    dtype, op for op as the reference writes it; the suppression is libfdms."""
    src_device = prediction.device
    pred = _cuda(prediction)
    bs = pred.shape[0]
    empty = torch.zeros((0, 6), device=src_device)
    mode = IOU_MODE if iou_mode is None else iou_mode
    xc = pred[..., 4] > conf_thres
    per_image = []
    for xi in range(bs):
        x = pred[xi][xc[xi]]
        if labels and len(labels[xi]):
            lab = torch.as_tensor(labels[xi], device=x.device, dtype=x.dtype)
            v = torch.zeros((len(lab), nc + 5), device=x.device, dtype=x.dtype)
            v[:, :4] = lab[:, 1:5]
            v[:, 4] = 1.0
            v[range(len(lab)), lab[:, 0].long() + 5] = 1.0
            x = torch.cat((x, v), 0)
        if not x.shape[0]:
            per_image.append(None)
            continue
        x = x.clone()
        x[:, 5:] *= x[:, 4:5]
        box = xywh2xyxy(x[:, :4])
        if multi_label:
            i, j = (x[:, 5:] > conf_thres).nonzero(as_tuple=False).T
            x = torch.cat((box[i], x[i, j + 5, None], j[:, None].to(x.dtype)), 1)
        else:
            conf, j = x[:, 5:].max(1, keepdim=True)
            x = torch.cat((box, conf, j.to(x.dtype)), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes, device=x.device)).any(1)]
        if x.shape[0] > MAX_NMS:
            x = x[x[:, 4].argsort(descending=True)[:MAX_NMS]]
        per_image.append(x if x.shape[0] else None)
    cap = max([x.shape[0] for x in per_image if x is not None], default=0)
    if cap == 0:
        return [empty] * bs
    rows = torch.zeros((bs, cap, 5), dtype=torch.float32, device=pred.device)
    count = torch.zeros((bs,), dtype=torch.int32, device=pred.device)
    for xi, x in enumerate(per_image):
        if x is None:
            continue
        c = x[:, 5:6] * (0 if agnostic else 4096)  # :582, in the prediction's dtype like the reference
        rows[xi, :x.shape[0], :4] = (x[:, :4] + c).float()
        rows[xi, :x.shape[0], 4] = x[:, 4].float()
        count[xi] = x.shape[0]
    keep, n_keep = ops.nms_segments(rows, count, cap, iou_thres, mode, max_nms=0)
    nk = n_keep.cpu()
    if int(nk.min()) < 0:
        raise _lib.FdmsError("non_max_suppression: a segment exceeded the suppression-mask budget (raise FDMS_MASK_BUDGET_MB)")
    out = []
    for xi, x in enumerate(per_image):
        if x is None or int(nk[xi]) == 0:
            out.append(empty)
            continue
        out.append(x[keep[xi, :int(nk[xi])].long()].to(src_device))
    return out


def _append_labels(c, labels, n_lab, rows, conf_thres):
    """utils/general.py:533-541 + 547-563 for the apriori label rows [cls, x, y, w, h]: box = xywh2xyxy, obj = cls = 1
    so conf = 1.0, appended AFTER the image's filtered predictions (torchvision's stable sort keeps that order among
    equal scores: the tie key continues after the prediction rows).  Returns the capacity the longest list needs."""
    counts = c.count.cpu()  # host round trip: autolabelling is an offline tool of the reference, not the hot path
    need = 0
    for b, n in enumerate(n_lab):
        if not n:
            need = max(need, int(counts[b]))
            continue
        lab = torch.as_tensor(labels[b], dtype=torch.float32, device=c.cand.device).reshape(n, -1)
        if bool((lab[:, 0].long() != 0).any()):
            raise NotImplementedError("non_max_suppression: labels of class != 0 on a single-class head")
        if not (1.0 > conf_thres):  # conf = 1.0 > conf_thres (:563) fails: the label rows are dropped
            need = max(need, int(counts[b]))
            continue
        k0 = int(counts[b])
        need = max(need, k0 + n)
        if k0 + n > c.cap:
            continue
        v = torch.zeros((n, 6), dtype=torch.float32, device=c.cand.device)
        v[:, :4] = xywh2xyxy(lab[:, 1:5])
        v[:, 4] = 1.0
        c.cand[b, k0:k0 + n] = v
        c.row[b, k0:k0 + n] = torch.arange(rows, rows + n, dtype=torch.int32, device=c.cand.device)
        c.count[b] = k0 + n
    return need


def xywh2xyxy(x):  # utils/general.py:299 (plain tensor arithmetic; the fused kernels do this inside the filter)
    y = x.clone() if isinstance(x, torch.Tensor) else x.copy()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def _inplace_on_device(coords, fn):
    """Run fn on a CUDA fp32 tensor that aliases `coords` when possible, else copy there and back."""
    if coords.is_cuda and coords.dtype == torch.float32 and coords.dim() == 2 and coords.stride(1) == 1:
        fn(coords)
        return coords
    tmp = _cuda(coords).float().contiguous()
    fn(tmp)
    coords.copy_(tmp.to(coords.dtype))
    return coords


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None, kpt_label=False, step=2):
    """utils/general.py:374.  Mutates `coords` in place and returns it, including the reference's
    clip quirk (boxes: only ROWS 0..3 are clipped, :389).  `ratio_pad` = (gain or (gain, ...), (pad_x, pad_y))."""
    if ratio_pad is not None:  # :379-383 (test.py:172 passes the dataset's ((h/h0, w/w0), (dw, dh)))
        gain = ratio_pad[0]
        if isinstance(gain, (list, tuple)):
            gain = gain[0]
        ratio_pad = (float(gain), (float(ratio_pad[1][0]), float(ratio_pad[1][1])))
    if coords.shape[0] == 0:
        return coords
    kstep = int(step) if kpt_label else 0
    return _inplace_on_device(coords, lambda t: ops.scale_coords_(img1_shape, t, img0_shape, kstep, ratio_pad))


def clip_coords(boxes, img_shape, step=2):  # utils/general.py:401
    boxes[:, 0::step].clamp_(0, img_shape[1])
    boxes[:, 1::step].clamp_(0, img_shape[0])
