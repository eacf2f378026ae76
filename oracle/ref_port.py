"""Library-call restatement of the reference's CPU hot path (TEST / BASELINE INFRASTRUCTURE ONLY).

Used by `bench.py`'s reference arm ONLY when no reference tree can be found (neither `/root/reference` nor
`baseline/_ref/`): it then times the same third-party calls the reference makes -- `cv2.resize(INTER_LINEAR)`,
`cv2.copyMakeBorder`, ATen elementwise ops, `torchvision.ops.nms` -- in the same order and at the same granularity
(one image, one scale at a time), so the baseline keeps the speed of the reference's libraries instead of the speed of
the scalar C oracle.  Each function cites what it restates; `tests/test_oracle_golden.py` checks it against the
committed fixtures of the reference's outputs and `tests/ref_live_check.py` against the live reference.
"""
import cv2
import numpy as np
import torch
from torchvision.ops import nms as tv_nms


def letterbox(img, new_shape, color=(114, 114, 114), stride=32):
    """utils/datasets.py:873-903 with auto=True, scaleFill=False, scaleup=True."""
    h0, w0 = img.shape[:2]
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h0, new_shape[1] / w0)
    unpad = (int(round(w0 * r)), int(round(h0 * r)))
    dw, dh = np.mod(new_shape[1] - unpad[0], stride) / 2, np.mod(new_shape[0] - unpad[1], stride) / 2
    if (w0, h0) != unpad:
        img = cv2.resize(img, unpad, interpolation=cv2.INTER_LINEAR)
    t, b = int(round(dh - 0.1)), int(round(dh + 0.1))
    l, rr = int(round(dw - 0.1)), int(round(dw + 0.1))
    return cv2.copyMakeBorder(img, t, b, l, rr, cv2.BORDER_CONSTANT, value=color)


def pack(img):
    """multi_scale_face_detector.py:96-105 on a CPU device: BGR->RGB, CHW, float, /255, batch axis."""
    x = torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))).float()
    x /= 255.0
    return x.unsqueeze(0)


def decode(levels, strides, anchors_px):
    """models/yolo.py:273-308 after the head's convolutions, inference branch, inplace=True, nkpt=5."""
    z = []
    for i, raw in enumerate(levels):
        bs, _, ny, nx = raw.shape
        x = raw.view(bs, 3, 21, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        det, kpt = x[..., :6].clone(), x[..., 6:].clone()
        yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
        grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
        y = det.sigmoid()
        xy = (y[..., 0:2] * 2. - 0.5 + grid) * strides[i]
        wh = (y[..., 2:4] * 2) ** 2 * torch.as_tensor(anchors_px[i], dtype=torch.float32).view(1, 3, 1, 1, 2)
        kpt[..., 0::3] = (kpt[..., 0::3] * 2. - 0.5 + grid[..., 0:1].repeat(1, 1, 1, 1, 5)) * strides[i]
        kpt[..., 1::3] = (kpt[..., 1::3] * 2. - 0.5 + grid[..., 1:2].repeat(1, 1, 1, 1, 5)) * strides[i]
        kpt[..., 2::3] = kpt[..., 2::3].sigmoid()
        z.append(torch.cat((xy, wh, y[..., 4:], kpt), -1).view(bs, -1, 21))
    return torch.cat(z, 1)


def non_max_suppression(pred, conf_thres, iou_thres):
    """utils/general.py:506-600 for nc=1, kpt_label=5, no labels / classes / multi_label."""
    out = []
    for x in pred:
        x = x[x[..., 4] > conf_thres]
        if not x.shape[0]:
            out.append(torch.zeros((0, 6)))
            continue
        x[:, 5:6] *= x[:, 4:5]
        box = x[:, :4].clone()
        box[:, 0], box[:, 1] = x[:, 0] - x[:, 2] / 2, x[:, 1] - x[:, 3] / 2
        box[:, 2], box[:, 3] = x[:, 0] + x[:, 2] / 2, x[:, 1] + x[:, 3] / 2
        conf, j = x[:, 5:6].max(1, keepdim=True)
        x = torch.cat((box, conf, j.float(), x[:, 6:]), 1)[conf.view(-1) > conf_thres]
        if x.shape[0] > 30000:
            x = x[x[:, 4].argsort(descending=True)[:30000]]
        out.append(x[tv_nms(x[:, :4] + x[:, 5:6] * 4096, x[:, 4], iou_thres)] if x.shape[0] else torch.zeros((0, 6)))
    return out


def detect_image(image, maps, sizes, strides, anchors_px, conf_thres, iou_thres):
    """multi_scale_face_detector.py:110-166 + 203-288 for one in-memory image; maps = {(H, W): [levels (63, ny, nx)]}."""
    rows = []
    h0, w0 = image.shape[:2]
    for si, size in enumerate(sizes):
        x = pack(letterbox(image, size))
        lv = [torch.from_numpy(l[None]).float() for l in maps[(int(x.shape[2]), int(x.shape[3]))]]
        for d in non_max_suppression(decode(lv, strides, anchors_px), conf_thres, iou_thres):
            if len(d):
                k = max(h0, w0) / x.shape[2]                       # ppyp:122-157
                d[:, [0, 2]] = (d[:, [0, 2]] * k).clamp(0, w0)
                d[:, [1, 3]] = (d[:, [1, 3]] * k).clamp(0, h0)
                d[:, :4] = d[:, :4].round()
                n = d.numpy()[:, :6]
                rows.append(np.hstack((n, np.ones((len(n), 1)) * si)))
    if not rows:
        return np.array([])
    m = np.vstack(rows)
    area = (m[:, 2] - m[:, 0]) * (m[:, 3] - m[:, 1])              # msfd:168-201
    idx, S, w = m[:, 6].astype(int), len(sizes), np.ones(len(m))
    w[(area < 1024) & (idx >= S - 2)] *= 1.2
    w[(area >= 1024) & (area <= 16384) & (idx == S // 2)] *= 1.1
    w[(area > 16384) & (idx <= 1)] *= 1.2
    keep = tv_nms(torch.from_numpy(m[:, :4]).float(), torch.from_numpy(m[:, 4] * w).float(), iou_thres)
    return m[keep.numpy()]
