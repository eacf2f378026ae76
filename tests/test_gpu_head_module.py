"""The patched head against the REAL reference IKeypoint (models/yolo.py:215-313) with random weights: the
weights, inputs and outputs in tests/golden/ikeypoint.npz come from the reference's own module (ImplicitA -> 1x1
conv -> ImplicitM, keypoint conv, cat, decode); here a structurally identical module carries the same weights and
`yolo_head.patch_head` replaces everything after the convolutions with the libfdms decode.  GPU only."""
import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import general, ops, yolo_head

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _ImplicitA(torch.nn.Module):  # models/common.py:55-63: x + implicit
    def __init__(self, c):
        super().__init__()
        self.implicit = torch.nn.Parameter(torch.zeros(1, c, 1, 1))

    def forward(self, x):
        return self.implicit.expand_as(x) + x


class _ImplicitM(torch.nn.Module):  # models/common.py:66-74: x * implicit
    def __init__(self, c):
        super().__init__()
        self.implicit = torch.nn.Parameter(torch.ones(1, c, 1, 1))

    def forward(self, x):
        return self.implicit.expand_as(x) * x


class _Head(torch.nn.Module):
    """Attribute layout of models/yolo.py::IKeypoint (nc=1, nkpt=5, single-conv keypoint branch)."""

    def __init__(self, ch, anchors_px, strides):
        super().__init__()
        self.nc, self.nkpt, self.no_det, self.no_kpt, self.no, self.nl, self.na = 1, 5, 6, 15, 21, len(ch), 3
        a = torch.as_tensor(np.asarray(anchors_px, np.float32)).view(self.nl, -1, 2)
        self.register_buffer("anchors", a.clone())
        self.register_buffer("anchor_grid", a.clone().view(self.nl, 1, -1, 1, 1, 2))
        self.m = torch.nn.ModuleList(torch.nn.Conv2d(c, self.no_det * self.na, 1) for c in ch)
        self.ia = torch.nn.ModuleList(_ImplicitA(c) for c in ch)
        self.im = torch.nn.ModuleList(_ImplicitM(self.no_det * self.na) for _ in ch)
        self.m_kpt = torch.nn.ModuleList(torch.nn.Conv2d(c, self.no_kpt * self.na, 1) for c in ch)
        self.stride = torch.tensor([float(s) for s in strides])


def _load(golden):
    g = golden("ikeypoint")
    head = _Head((8, 16, 32), g["w_anchor_grid"].reshape(3, 3, 2), (8, 16, 32))
    sd = {k[2:].replace("__", "."): torch.from_numpy(g[k]) for k in g.files if k.startswith("w_")}
    missing, unexpected = head.load_state_dict(sd, strict=True)
    head.to(DEV).eval()
    return g, head


def test_patched_head_equals_reference_ikeypoint(golden):
    g, head = _load(golden)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yolo_head.patch_head(head)
    xs = [torch.from_numpy(g["x%d" % i]).to(DEV) for i in range(3)]
    with torch.no_grad():
        pred, raw = head([x.clone() for x in xs])
    assert pred.shape == g["pred"].shape
    # the 1x1 convolutions run in cuDNN (fp32, different summation order than the CPU reference): compare the
    # conv outputs loosely and the decode -- the part libfdms replaces -- tightly, on the SAME conv outputs
    for i, r in enumerate(raw):
        assert r.shape == g["raw%d" % i].shape
        assert np.max(np.abs(r.cpu().numpy() - g["raw%d" % i])) < 1e-4
    from oracle import oracle as O
    raw_nchw = [r.permute(0, 1, 4, 2, 3).reshape(r.shape[0], 63, r.shape[2], r.shape[3]).contiguous().cpu().numpy() for r in raw]
    ref = O.decode_head(raw_nchw, (8, 16, 32), g["w_anchor_grid"].reshape(3, 3, 2))
    assert np.max(np.abs(pred.cpu().numpy() - ref)) <= 1e-4
    # and end to end against the reference module's own output (conv rounding amplified by stride / anchors)
    assert np.max(np.abs(pred.cpu().numpy() - g["pred"])) <= 2e-2
    assert np.mean(np.abs(pred.cpu().numpy() - g["pred"]) <= 1e-3) > 0.999


def test_reference_pred_through_nms(golden):
    """non_max_suppression drop-in on the reference module's own decoded output: bit-exact rows."""
    g, _ = _load(golden)
    pred = torch.from_numpy(g["pred"]).to(DEV)
    res = general.non_max_suppression(pred, float(g["thr"]), 0.45, iou_mode=ops.IOU_TV_CPU)
    for b in range(2):
        assert np.array_equal(res[b].cpu().numpy(), g["nms_b%d" % b])
