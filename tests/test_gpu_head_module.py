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


# ---- every other head variant the reference can run (tests/golden/heads.npz, make_golden.py::gen_heads) ------------
HEAD_VARIANTS = {  # name: class of the reference module
    "detect_nc3": "Detect", "detect_nc80": "Detect", "idetect_nc1": "IDetect", "idetect_pose": "IDetect",
    "detect_face": "Detect", "ikeypoint_k3": "IKeypoint",
}


def _variant(golden, name):
    g = golden("heads")
    nc, nkpt, na, inplace = (int(v) for v in g[name + "__meta"])
    return g, nc, nkpt, na, bool(inplace), 5 + nc + 3 * nkpt


@pytest.mark.parametrize("name", list(HEAD_VARIANTS))
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_decode_dense_head_vs_oracle_and_reference(golden, name, dtype):
    from oracle import oracle as O
    g, nc, nkpt, na, _, no = _variant(golden, name)
    raw = [g["%s__raw%d" % (name, i)].astype(dtype) for i in range(3)]
    anchors = g[name + "__anchors_px"]
    pred = ops.decode_dense_head([torch.from_numpy(r).to(DEV) for r in raw], (8, 16, 32), anchors, na, no, nkpt).cpu().numpy()
    ref = O.decode_head_general(raw, (8, 16, 32), anchors, na, no, nkpt)
    assert pred.shape == ref.shape == g[name + "__pred"].shape
    assert np.array_equal(pred, ref)  # same correctly rounded sigmoid, every operation rounded separately: bit identical
    if dtype is np.float32:  # the reference module's own output (ATen CPU exp: <= 1 ulp in the sigmoid, see test_oracle_golden)
        gp = g[name + "__pred"]
        assert np.all(np.abs(pred - gp) <= np.maximum(1e-4, 4 * np.spacing(np.abs(gp))))


def test_decode_dense_head_equals_face_kernel(golden):
    """On the face layout (na=3, no=21, nkpt=5) the general kernel and the specialised one are the same function."""
    g = golden("heads")
    raw = [torch.from_numpy(g["detect_face__raw%d" % i]).to(DEV) for i in range(3)]
    a = ops.decode_dense(raw, (8, 16, 32), g["detect_face__anchors_px"])
    b = ops.decode_dense_head(raw, (8, 16, 32), g["detect_face__anchors_px"], 3, 21, 5)
    assert torch.equal(a, b)


def test_decode_dense_head_ragged_tail_and_wide_rows():
    """Cell counts that are no multiple of the CTA size, one anchor, the widest supported row (staging in 32-cell CTAs)."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    for na, no, nkpt, hw in ((1, 6, 0, (7, 13)), (8, 9, 1, (5, 3)), (2, 383, 0, (3, 11)), (3, 96, 30, (9, 17)), (5, 200, 0, (6, 7))):
        raw = [rng.normal(0, 2, size=(2, na * no, hw[0] * k, hw[1] * k)).astype(np.float32) for k in (2, 1)]
        anchors = rng.uniform(4, 300, size=(2, na, 2)).astype(np.float32)
        pred = ops.decode_dense_head([torch.from_numpy(r).to(DEV) for r in raw], (8, 16), anchors, na, no, nkpt).cpu().numpy()
        assert np.array_equal(pred, O.decode_head_general(raw, (8, 16), anchors, na, no, nkpt)), (na, no, nkpt)


def test_decode_dense_head_rejects_bad_arguments():
    x = [torch.zeros((1, 3 * 22, 4, 4), device=DEV)]
    with pytest.raises(ValueError):
        ops.decode_dense_head(x, [8], np.ones((1, 3, 2)), 3, 22, 5)  # a keypoint head has 6 + 3*nkpt columns
    with pytest.raises(NotImplementedError):
        ops.decode_dense_head(x, [8], np.ones((1, 9, 2)), 9, 22, 0)  # more anchors than FDMS_MAX_ANCHORS
    with pytest.raises(ValueError):
        ops.decode_dense_head(x, [8], np.ones((1, 3, 2)), 3, 21, 5)  # channel count does not match na * no


class _AnyHead(torch.nn.Module):
    """Attribute layout of models/yolo.py::Detect / IDetect / IKeypoint for any nc / nkpt / na (single-conv keypoint branch)."""

    def __init__(self, kind, ch, nc, nkpt, na, inplace, anchors_px, strides):
        super().__init__()
        self.nc, self.nkpt, self.no_det, self.no_kpt, self.nl, self.na = nc, nkpt, nc + 5, 3 * nkpt, len(ch), na
        self.no, self.inplace = self.no_det + self.no_kpt, inplace
        a = torch.as_tensor(np.asarray(anchors_px, np.float32)).view(self.nl, -1, 2)
        self.register_buffer("anchors", a.clone())
        self.register_buffer("anchor_grid", a.clone().view(self.nl, 1, -1, 1, 1, 2))
        self.m = torch.nn.ModuleList(torch.nn.Conv2d(c, self.no_det * na, 1) for c in ch)
        if kind != "Detect":
            self.ia = torch.nn.ModuleList(_ImplicitA(c) for c in ch)
            self.im = torch.nn.ModuleList(_ImplicitM(self.no_det * na) for _ in ch)
        self.m_kpt = torch.nn.ModuleList(torch.nn.Conv2d(c, self.no_kpt * na, 1) for c in ch)
        self.stride = torch.tensor([float(s) for s in strides])


# the test classes carry the reference's class names: patch_head keys the in-place keypoint write-back on them
class Detect(_AnyHead):
    pass


class IDetect(_AnyHead):
    pass


class IKeypoint(_AnyHead):
    pass


@pytest.mark.parametrize("name", [n for n in HEAD_VARIANTS if n != "detect_nc80"])
def test_patched_head_variants_equal_reference_modules(golden, name):
    g, nc, nkpt, na, inplace, no = _variant(golden, name)
    kind = HEAD_VARIANTS[name]
    head = {"Detect": Detect, "IDetect": IDetect, "IKeypoint": IKeypoint}[kind](
        kind, (8, 16, 32), nc, nkpt, na, inplace, g[name + "__anchors_px"], (8, 16, 32))
    pre = name + "__w_"
    sd = {k[len(pre):].replace("__", "."): torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)}
    head.load_state_dict(sd, strict=True)
    head.to(DEV).eval()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yolo_head.patch_head(head)
    xs = [torch.from_numpy(g["%s__x%d" % (name, i)]).to(DEV) for i in range(3)]
    with torch.no_grad():
        pred, xo = head([x.clone() for x in xs])
    gp = g[name + "__pred"]
    assert pred.shape == gp.shape
    # the convolutions run in cuDNN (other summation order than the CPU reference); their error is amplified by the
    # stride / the anchors, so the module-level comparison is loose and the decode itself is pinned above
    d = np.abs(pred.cpu().numpy() - gp)
    assert np.max(d / np.maximum(1.0, np.abs(gp))) <= 2e-3
    assert np.mean(d <= 1e-3 * np.maximum(1.0, np.abs(gp))) > 0.999
    mutated = nkpt > 0 and kind in ("Detect", "IDetect")
    off = 0
    for i, r in enumerate(xo):
        bs, na_, ny, nx, no_ = r.shape
        assert (na_, no_) == (na, no)
        raw_view = head.fdms_raw[i].view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2)
        if mutated:  # models/yolo.py:76-81 / :169-183: the keypoint columns of the returned tensor hold decoded values
            assert torch.equal(r[..., :6], raw_view[..., :6])
            assert torch.equal(r[..., 6:], pred[:, off:off + na * ny * nx].view(bs, na, ny, nx, no)[..., 6:])
            gx = g["%s__xo%d" % (name, i)]
            assert np.max(np.abs(r.cpu().numpy() - gx) / np.maximum(1.0, np.abs(gx))) <= 2e-3
        else:
            assert torch.equal(r, raw_view)
        off += na * ny * nx


def test_patch_head_refuses_what_the_reference_cannot_run():
    mk = lambda kind, nc, nkpt, inplace: {"Detect": Detect, "IDetect": IDetect}[kind](  # noqa: E731
        kind, (8,), nc, nkpt, 3, inplace, np.ones((1, 3, 2)), (8,))
    for args in (("Detect", 2, 5, True), ("IDetect", 1, 5, False), ("Detect", 80, 0, True)):
        with pytest.raises(NotImplementedError):
            yolo_head.patch_head(mk(*args))


def test_plain_head_through_multiclass_nms(golden):
    """A multi-class head end to end at the compat level: libfdms decode -> general.non_max_suppression (class offsets,
    utils/general.py:533-582) against the oracle's decode pushed through the same function."""
    from oracle import oracle as O
    g, nc, nkpt, na, _, no = _variant(golden, "detect_nc3")
    raw = [g["detect_nc3__raw%d" % i] for i in range(3)]
    anchors = g["detect_nc3__anchors_px"]
    pred = ops.decode_dense_head([torch.from_numpy(r).to(DEV) for r in raw], (8, 16, 32), anchors, na, no, nkpt)
    ref = torch.from_numpy(O.decode_head_general(raw, (8, 16, 32), anchors, na, no, nkpt)).to(DEV)
    a = general.non_max_suppression(pred, 0.05, 0.45, kpt_label=0)
    b = general.non_max_suppression(ref, 0.05, 0.45, kpt_label=0)
    assert sum(len(r) for r in a) > 10
    for ra, rb in zip(a, b):
        assert ra.shape[1] == 6 and torch.equal(ra, rb)


# ---- the half-model arithmetic of FDMS_NUMERICS=aten_cuda on fp16 maps (the reference's default mode on a GPU) --------------
def _aten_cuda(fn):
    from face_detection_multi_scale_b200 import _lib
    c = _lib.ctx(torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    c.set_numerics("aten_cuda")
    try:
        return fn()
    finally:
        c.set_numerics("exact")


def _half_close(got, ref, stride_max=32.0):
    """Equal except where libdevice's expf and numpy's differ in the last float bit AND that flips a half rounding:
    such an element moves by one half ulp (xy: one half ulp of the sigmoid times 2 * stride; wh = (2 s)^2 * anchor doubles
    the relative step, so two half ulps are allowed)."""
    same = got == ref
    assert same.mean() > 0.999, float(same.mean())
    d = np.abs(got - ref)[~same]
    r = np.abs(ref)[~same]
    assert np.all(d <= np.maximum(2.0 ** -9 * r, 2.0 ** -9 * stride_max)), float(d.max()) if d.size else 0.0  # <= 2 half ulps


@pytest.mark.parametrize("name", ["detect_face", "idetect_pose", "ikeypoint_k3", "idetect_nc1"])
def test_half_model_decode_matches_oracle(golden, name):
    """The dense decode kernels with aten_cuda numerics on fp16 maps round where ATen rounds for a half model
    (oracle.decode_head_half, pinned bit for bit against the reference modules on half tensors by tests/ref_live_check.py;
    the GPU pin against the real reference is tests/ref_gpu_check.py P1h / C / D)."""
    from oracle import oracle as O
    g, nc, nkpt, na, _, no = _variant(golden, name)
    raw = [g["%s__raw%d" % (name, i)].astype(np.float16) for i in range(3)]
    anchors = g[name + "__anchors_px"]
    ref = O.decode_head_half(raw, (8, 16, 32), anchors, na, no, nkpt)
    dev = [torch.from_numpy(r).to(DEV) for r in raw]
    got = _aten_cuda(lambda: ops.decode_dense_head(dev, (8, 16, 32), anchors, na, no, nkpt)).cpu().numpy()
    _half_close(got, ref)
    if na == 3 and no == 21:  # the specialised face kernel runs the same arithmetic
        got2 = _aten_cuda(lambda: ops.decode_dense(dev, (8, 16, 32), anchors)).cpu().numpy()
        assert np.array_equal(got, got2)
    # fp32 maps under aten_cuda keep the fp32 arithmetic (a float model): no half rounding
    raw32 = [torch.from_numpy(r.astype(np.float32)).to(DEV) for r in raw]
    a = _aten_cuda(lambda: ops.decode_dense_head(raw32, (8, 16, 32), anchors, na, no, nkpt)).cpu().numpy()
    assert np.max(np.abs(a - O.decode_head_general([r.astype(np.float32) for r in raw], (8, 16, 32), anchors, na, no, nkpt))) <= 1e-3
    assert not np.array_equal(a, got)


def test_half_model_fused_filter_matches_dense_decode(golden):
    """The fused decode + filter under the half-model arithmetic keeps exactly the rows the dense half decode passes through
    the confidence test (conf = cls * obj on the half-rounded columns, float32 product), with identical candidate rows."""
    g = golden("heads")
    raw = [torch.from_numpy(g["detect_face__raw%d" % i].astype(np.float16)).to(DEV) for i in range(3)]
    anchors = g["detect_face__anchors_px"]
    for thr in (0.05, 0.25, 0.5):
        def run():
            pred = ops.decode_dense(raw, (8, 16, 32), anchors)
            c = ops.decode_filter([raw], (8, 16, 32), anchors, thr, cap=4096)
            return pred, c
        pred, c = _aten_cuda(run)
        pred = pred.cpu().numpy()
        for b in range(pred.shape[0]):
            p = pred[b]
            conf = (p[:, 5] * p[:, 4]).astype(np.float32)
            want = np.nonzero((p[:, 4] > np.float32(thr)) & (conf > np.float32(thr)))[0]
            n = int(c.count[b].item())
            rows = c.row[b, :n].cpu().numpy()
            order = np.argsort(rows)
            assert np.array_equal(rows[order], want), (thr, b, n, len(want))
            cand = c.cand[b, :n].cpu().numpy()[order]
            q = p[want]
            box = np.stack([q[:, 0] - q[:, 2] / 2, q[:, 1] - q[:, 3] / 2, q[:, 0] + q[:, 2] / 2, q[:, 1] + q[:, 3] / 2], 1)
            assert np.array_equal(cand[:, :4], box.astype(np.float32)) and np.array_equal(cand[:, 4], conf[want])
            assert np.array_equal(cand[:, 6:], q[:, 6:])


@pytest.mark.parametrize("name", ("detect_face", "idetect_pose", "ikeypoint_k3", "idetect_nc1"))
def test_half_model_decode_matches_reference_fixture(golden, name):
    """The same kernels against what the reference's own modules return for a half model (tests/golden/heads_half.npz)."""
    g = golden("heads_half")
    nc, nkpt, na, _ = (int(v) for v in g[name + "__meta"])
    raw = [torch.from_numpy(g["%s__raw%d" % (name, i)]).to(DEV) for i in range(3)]
    got = _aten_cuda(lambda: ops.decode_dense_head(raw, (8, 16, 32), g[name + "__anchors_px"], na, 5 + nc + 3 * nkpt, nkpt))
    assert got.dtype == torch.float32
    _half_close(got.cpu().numpy(), g[name + "__pred"])


def test_half_model_landmarks_at_emit_equal_the_staged_path():
    """fdms_postprocess decodes the landmarks of the FINAL detections straight from the head maps (merge_rows.cuh); under the
    half-model arithmetic that must equal the staged path -- decode_filter with all 21 columns, NMS, merge taking the
    landmark columns from the candidate rows -- bit for bit, detections included."""
    from workloads import standin, synth
    h, w, sizes, B = 270, 480, [160, 256, 320], 2
    _, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=25, seed=9, dtype=np.float16)
    sl = [[torch.from_numpy(l).to(DEV) for l in heads[s]] for s in range(len(sizes))]
    in_h = [s[0] for s in shapes]

    def run():
        plan = ops.PostPlan(sl, synth.P5_STRIDES, synth.P5_ANCHORS, in_h, (h, w), 0.4, 0.5, cap=2048, out_cap=2048,
                            iou_mode=ops.IOU_TV_CPU, with_kpts=True)
        a = plan.run()
        c = ops.decode_filter(sl, synth.P5_STRIDES, synth.P5_ANCHORS, 0.4, cap=2048, landmarks=True)
        ops.nms_segments(c.cand, c.count, c.cap, 0.5, ops.IOU_TV_CPU, tie=c.row, keep=c.keep, n_keep=c.n_keep)
        b = ops.merge_scales(c, B, len(sizes), in_h, (h, w), 0.5, 2048, with_kpts=True)
        torch.cuda.synchronize()
        return a, b
    a, b = _aten_cuda(run)
    assert torch.equal(a.n_out, b.n_out) and int(a.n_out.sum()) > 10
    for i in range(B):
        n = int(a.n_out[i])
        assert torch.equal(a.dets[i, :n], b.dets[i, :n])
        assert torch.equal(a.kpts[i, :n], b.kpts[i, :n])
    # and the half-model landmarks differ from the fp32 arithmetic on the same maps (the mode is really on)
    a32 = ops.PostPlan(sl, synth.P5_STRIDES, synth.P5_ANCHORS, in_h, (h, w), 0.4, 0.5, cap=2048, out_cap=2048,
                       iou_mode=ops.IOU_TV_CPU, with_kpts=True).run()
    torch.cuda.synchronize()
    n0 = min(int(a.n_out[0]), int(a32.n_out[0]))
    assert not torch.equal(a.kpts[0, :n0], a32.kpts[0, :n0])


def test_half_model_detector_matches_oracle_pipeline():
    """The whole fused path (decode + filter, NMS, rescale, cross-scale merge) of a half model under aten_cuda numerics
    against the oracle's half-model pipeline (oracle.detect_from_heads(half_model=True), pinned against the live reference's
    detect_multi_scale on half tensors by tests/ref_live_check.py): the same (k, 7) arrays."""
    from oracle import oracle as O
    from workloads import standin, synth
    from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector
    h, w, sizes, B = 270, 480, [160, 256, 320], 3
    images, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=25, seed=3, dtype=np.float16)
    model = standin.SyntheticHeadModel()
    for s, hw in enumerate(shapes):
        model.set_heads(hw, heads[s], pin=True)
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU, half=True)
    out = _aten_cuda(lambda: det.detect_multi_scale_batch(images))
    n_total = 0
    for b in range(B):
        ref, _ = O.detect_from_heads([[l[b:b + 1] for l in heads[s]] for s in range(len(sizes))], shapes, (h, w, 3),
                                     synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5, O.TV_CPU, half_model=True)
        assert out[b].shape == ref.shape and np.array_equal(out[b], ref), b
        n_total += len(ref)
    assert n_total > 20
