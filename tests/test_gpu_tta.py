"""Test-time augmentation around the backbone on libfdms: scale_img (utils/torch_utils.py:247-257),
_descale_pred and the concatenation of Model.forward_augment (models/yolo.py:363-417).  Fixture tests/golden/tta.npz
comes from the reference's own functions (make_golden.py tta)."""
import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import torch_utils as TU
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2.4e-7  # fp32, inputs in [0, 1]: 2 ulp at 1.0 (ATen contracts the 4-tap sum differently)


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_scale_img_matches_reference_golden(golden):
    g = golden("tta")
    x = gpu(g["x"])
    for i in range(5):
        ratio, flip, gs = g["s%d_args" % i]
        got = TU.scale_img(x, float(ratio), gs=int(gs), flip=int(flip) or None).cpu().numpy()
        ref = g["s%d" % i]
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= TOL
        # the oracle restates the kernel's own operation order: bit exact
        assert np.array_equal(got, O.scale_img(g["x"], float(ratio), gs=int(gs), flip=int(flip) or None))
    got = TU.scale_img(x, 0.67, same_shape=True).cpu().numpy()
    assert got.shape == g["same_shape"].shape and np.max(np.abs(got - g["same_shape"])) <= TOL


@pytest.mark.parametrize("shape,ratio,flip", [((1, 3, 640, 640), 0.83, 3), ((3, 3, 384, 672), 0.67, None), ((2, 1, 37, 53), 0.5, 2)])
def test_scale_img_vs_oracle_and_torch(shape, ratio, flip):
    rng = np.random.default_rng(1)
    x = rng.random(shape, dtype=np.float32)
    got = TU.scale_img(gpu(x), ratio, gs=32, flip=flip)
    assert np.array_equal(got.cpu().numpy(), O.scale_img(x, ratio, gs=32, flip=flip))
    # and against live PyTorch on the same device (the arithmetic the reference executes on a GPU)
    xt = gpu(x).flip(flip) if flip else gpu(x)
    rs = (int(shape[2] * ratio), int(shape[3] * ratio))
    ref = torch.nn.functional.interpolate(xt, size=rs, mode="bilinear", align_corners=False)
    assert torch.max(torch.abs(got[:, :, :rs[0], :rs[1]] - ref)).item() <= TOL
    assert torch.all(got[:, :, rs[0]:, :] == 0.447) and torch.all(got[:, :, :, rs[1]:] == 0.447)
    # FDMS_NUMERICS=aten_cuda: the FMA contraction of ATen's CUDA kernel (csrc/tta.cu) -> bit-identical to F.interpolate here
    from face_detection_multi_scale_b200 import _lib
    c = _lib.ctx(torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    c.set_numerics("aten_cuda")
    try:
        got_a = TU.scale_img(gpu(x), ratio, gs=32, flip=flip)
        got_h = TU.scale_img(gpu(x).half(), ratio, gs=32, flip=flip)
    finally:
        c.set_numerics("exact")
    assert torch.equal(got_a[:, :, :rs[0], :rs[1]], ref)
    ref_h = torch.nn.functional.interpolate(xt.half(), size=rs, mode="bilinear", align_corners=False)
    assert torch.equal(got_h[:, :, :rs[0], :rs[1]], ref_h)


def test_scale_img_fp16():
    rng = np.random.default_rng(2)
    x = rng.random((2, 3, 96, 160), dtype=np.float32).astype(np.float16)
    got = TU.scale_img(gpu(x), 0.83, gs=32, flip=3).cpu().numpy()
    ref = O.scale_img(x.astype(np.float32), 0.83, gs=32, flip=3).astype(np.float16)  # fp32 arithmetic, one rounding
    assert got.dtype == np.float16 and np.array_equal(got, ref)


def test_descale_pred_and_cat_golden(golden):
    g = golden("tta")
    p = gpu(g["p"])
    out = torch.empty((2, 120, 21), dtype=torch.float32, device=DEV)
    r0 = 0
    for i, (si, fi) in enumerate(zip(TU.TTA_SCALES, TU.TTA_FLIPS)):
        one = TU.descale_pred(p.clone(), fi, si, (48, 80))
        assert np.array_equal(one.cpu().numpy(), g["d%d" % i])
        TU.descale_pred(p, fi, si, (48, 80), out=out, row_offset=r0)
        r0 += 40
    assert np.array_equal(out.cpu().numpy(), g["cat"])
    assert np.array_equal(TU.descale_pred(p.clone(), 2, 0.83, (48, 80)).cpu().numpy(), g["d_ud"])


def test_forward_augment_structure():
    """forward_augment with a stand-in forward_once: the three passes see scale_img(x.flip(f), s) and their
    predictions come back de-scaled and concatenated in pass order (models/yolo.py:363-374)."""
    rng = np.random.default_rng(4)
    x = rng.random((2, 3, 64, 96), dtype=np.float32)
    seen = []

    def forward_once(xi):
        seen.append(xi.cpu().numpy())
        bs = xi.shape[0]
        rows = xi.shape[2] * xi.shape[3] // 64
        pred = xi.reshape(bs, -1)[:, :rows * 21].reshape(bs, rows, 21) * 100.0
        return pred, None

    out, aux = TU.forward_augment(forward_once, gpu(x), gs=32)
    assert aux is None
    parts = []
    for k, (si, fi) in enumerate(zip(TU.TTA_SCALES, TU.TTA_FLIPS)):
        xi = O.scale_img(x, si, gs=32, flip=fi)
        assert np.array_equal(seen[k], xi)
        rows = xi.shape[2] * xi.shape[3] // 64
        pred = xi.reshape(2, -1)[:, :rows * 21].reshape(2, rows, 21) * np.float32(100.0)
        parts.append(O.descale_pred(pred, fi, si, (64, 96)))
    assert np.array_equal(out.cpu().numpy(), np.concatenate(parts, 1))
