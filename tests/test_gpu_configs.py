"""BASELINE.json configs as parity cases at (or near) their full sizes, plus size-independent properties.
GPU only.  cfg1/cfg2 shapes are covered in test_gpu_parity.py (letterbox digests at 1080p/640..1280)."""
import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import _lib, ops
from workloads import standin, synth
from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _oracle_detect(heads, shapes, b, hw, strides, anchors, conf, iou):
    return O.detect_from_heads([[l[b:b + 1] for l in heads[s]] for s in range(len(shapes))], shapes, (hw[0], hw[1], 3),
                               strides, anchors, conf, iou, O.TV_CPU)


def test_cfg3_dense_crowd_20k_candidates():
    """configs[2]: ~20k pre-NMS candidates per image over 3 scales, landmarks on."""
    h, w, sizes = 1080, 1920, [640, 960, 1280]
    images, heads, shapes = standin.make_batch(h, w, sizes, 2, n_faces=1300, seed=11, dtype=np.float16)
    sl = [[gpu(l) for l in heads[s]] for s in range(3)]
    plan = ops.PostPlan(sl, synth.P5_STRIDES, synth.P5_ANCHORS, [s[0] for s in shapes], (h, w), 0.5, 0.5, cap=16384,
                        out_cap=16384, iou_mode=ops.IOU_TV_CPU, with_kpts=True)
    out = plan.run()
    torch.cuda.synchronize()
    cnt = plan.cands.count.cpu().numpy().reshape(2, 3)
    assert 15000 < cnt[0].sum() < 30000, cnt  # the dense-crowd operating point
    for b in range(2):
        ref, per_scale = _oracle_detect(heads, shapes, b, (h, w), synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5)
        n = int(out.n_out[b])
        assert n == len(ref)
        assert np.array_equal(out.dets[b, :n].cpu().numpy().astype(np.float64), ref)
        assert [int(x) for x in plan.cands.n_keep.cpu().numpy().reshape(2, 3)[b]] == [len(p) for p in per_scale]
        k = out.kpts[b, :n].cpu().numpy()
        assert np.all(k[:, 0::3] >= 0) and np.all(k[:, 0::3] <= w) and np.all(k[:, 1::3] <= h)


def test_cfg4_w6_p6_head_4_scales_4k():
    """configs[3]: yolov7-w6 P6 head (strides 8/16/32/64), 4-scale pyramid up to 1536, 4K source."""
    h, w, sizes = 2160, 3840, [640, 960, 1280, 1536]
    images, heads, shapes = standin.make_batch(h, w, sizes, 1, n_faces=50, seed=23, dtype=np.float16,
                                               strides=synth.P6_STRIDES, anchors_px=synth.P6_ANCHORS, max_size=500.0)
    assert [tuple(s) for s in shapes] == [(384, 640), (576, 960), (768, 1280), (896, 1536)]
    geoms = [_lib.letterbox_geometry(h, w, s, auto=True, stride=64) for s in sizes]
    outs = ops.pyramid(gpu(images), geoms, out_dtype=torch.float16)
    for s, size in enumerate(sizes):
        lb, _, _ = O.letterbox(images[0], size, auto=True, stride=64)
        ref = O.pack_chw(lb, flip=True, dtype=np.float16)
        assert np.array_equal(outs[s][0].cpu().numpy().view(np.uint16), ref.view(np.uint16)), size
    sl = [[gpu(l) for l in heads[s]] for s in range(4)]
    plan = ops.PostPlan(sl, synth.P6_STRIDES, synth.P6_ANCHORS, [s[0] for s in shapes], (h, w), 0.5, 0.5, cap=4096,
                        out_cap=4096, iou_mode=ops.IOU_TV_CPU)
    out = plan.run()
    ref, _ = _oracle_detect(heads, shapes, 0, (h, w), synth.P6_STRIDES, synth.P6_ANCHORS, 0.5, 0.5)
    n = int(out.n_out[0])
    assert n == len(ref) > 10
    assert np.array_equal(out.dets[0, :n].cpu().numpy().astype(np.float64), ref)
    assert set(np.unique(ref[:, 6]).astype(int)) <= {0, 1, 2, 3}


def test_cfg5_batch_single_scale_sharded_consistency():
    """configs[4]: large single-scale batch.  Any contiguous shard of the batch gives the same per-image result
    as the whole batch (what image sharding across GPUs relies on), and spot images equal the oracle."""
    h, w, sizes, B = 1080, 1920, [640], 24
    images, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=30, seed=31, dtype=np.float16)

    def run(lo, hi):
        model = standin.SyntheticHeadModel()
        model.set_heads(shapes[0], [l[lo:hi] for l in heads[0]], pin=True)
        det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU)
        return det.detect_multi_scale_batch(images[lo:hi])

    whole = run(0, B)
    parts = run(0, 12) + run(12, B)
    for a, b in zip(whole, parts):
        assert a.shape == b.shape and np.array_equal(a, b)
    for b in (0, 7, B - 1):
        ref, _ = _oracle_detect(heads, shapes, b, (h, w), synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5)
        assert whole[b].shape == ref.shape and np.array_equal(whole[b], ref)


def test_pyramid_properties_full_size():
    """Size-independent checks at cfg2's full size: border value, constant image stays constant, batch entries are
    independent, letterbox of an already-letterboxed shape is the identity."""
    h, w = 1080, 1920
    geoms = [_lib.letterbox_geometry(h, w, s, auto=True, stride=32) for s in (640, 960, 1280)]
    const = np.full((2, h, w, 3), 200, np.uint8)
    const[1] = synth.synth_image(h, w, 5)
    outs = ops.pyramid(gpu(const), geoms, out_dtype=torch.float16)
    v200, v114 = np.float16(np.float32(200) / np.float32(255)), np.float16(np.float32(114) / np.float32(255))
    for g, o in zip(geoms, outs):
        a = o[0].cpu().numpy()
        assert np.all(a[:, g.top:g.top + g.unpad_h, g.left:g.left + g.unpad_w] == v200)
        assert np.all(a[:, :g.top] == v114) and np.all(a[:, g.top + g.unpad_h:] == v114)
    alone = ops.pyramid(gpu(const[1:2]), geoms, out_dtype=torch.float16)
    for o, a in zip(outs, alone):
        assert torch.equal(o[1], a[0])
    img = synth.synth_image(384, 640, 9)
    g = _lib.letterbox_geometry(384, 640, 640, auto=True, stride=32)
    assert (g.dst_h, g.dst_w, g.resize) == (384, 640, 0)
    (same,) = ops.pyramid(gpu(img[None]), [g], out_dtype=torch.uint8, layout="hwc", flip_rb=False)
    assert np.array_equal(same[0].cpu().numpy(), img)


def test_nms_idempotent_and_sorted_large():
    """Properties at max_nms scale: output scores are non-increasing, kept boxes do not suppress each other, and
    NMS of the kept set keeps everything."""
    n = 30000
    p = synth.synth_pred(n, 800, 60, 77, extent=(3000.0, 3000.0))
    cand, _ = O.filter_xyxy(p, -1.0)
    rows = gpu(cand[None])
    count = torch.tensor([n], dtype=torch.int32, device=DEV)
    keep, n_keep = ops.nms_segments(rows, count, n, 0.45, ops.IOU_TV_CUDA)
    k = keep[0, :int(n_keep[0])].long()
    kept = rows[0, k]
    s = kept[:, 4].cpu().numpy()
    assert np.all(s[:-1] >= s[1:]) and len(s) > 500
    again = ops.nms(kept[:, :4].contiguous(), kept[:, 4].contiguous(), 0.45, ops.IOU_TV_CUDA)
    assert torch.equal(again, torch.arange(len(s), device=DEV))


def test_repeated_runs_are_bit_identical():
    """The whole path 40 times on the same inputs (cfg2 shapes, batch 8) with other work in flight on a second stream:
    pyramid planes, per-scale keep counts and merged detections must not change from run to run.  Candidate ORDER in
    the scratch lists is free (atomics); everything observable is not.  Guards the barrier-free parts of the kernels
    (mbarrier staging, ballot resolver, warp-aggregated atomics) against races -- compute-sanitizer is not available
    on this GPU pool."""
    h, w, sizes, bs = 1080, 1920, [640, 960, 1280], 8
    images, heads, shapes = standin.make_batch(h, w, sizes, bs, n_faces=60, seed=23, dtype=np.float16)
    img = gpu(images)
    geoms = [_lib.letterbox_geometry(h, w, s, auto=True, stride=32) for s in sizes]
    sl = [[gpu(l) for l in heads[s]] for s in range(3)]
    plan = ops.PostPlan(sl, synth.P5_STRIDES, synth.P5_ANCHORS, [s[0] for s in shapes], (h, w), 0.5, 0.5, cap=1024,
                        out_cap=1024, iou_mode=ops.IOU_TV_CUDA)
    side = torch.cuda.Stream()
    noise = torch.empty((64 << 20,), dtype=torch.uint8, device=DEV)
    first = None
    for it in range(40):
        with torch.cuda.stream(side):  # unrelated traffic perturbs the scheduling from run to run
            noise.fill_(it & 0xff)
        outs = ops.pyramid(img, geoms, out_dtype=torch.float16)
        out = plan.run()
        torch.cuda.synchronize()
        n_out = out.n_out.cpu().numpy().copy()
        dets = [out.dets[b, :n_out[b]].cpu().numpy().copy() for b in range(bs)]
        sig = ([o.view(torch.int16).to(torch.int64).sum().item() for o in outs], plan.cands.count.cpu().numpy().copy(),
               plan.cands.n_keep.cpu().numpy().copy(), n_out)
        if first is None:
            first = (sig, dets, [o.clone() for o in outs])
            continue
        for a, b_ in zip(sig, first[0]):
            assert np.array_equal(np.asarray(a), np.asarray(b_)), it
        for a, b_ in zip(dets, first[1]):
            assert np.array_equal(a, b_), it
        if it % 13 == 0:
            for a, b_ in zip(outs, first[2]):
                assert torch.equal(a, b_), it


def test_speculative_small_chain_falls_back_on_dense_batches():
    """The detector launches only the single-CTA kernels while its batches stay sparse (FDMS_POST_SMALL_ONLY) and
    verifies the guess from the exact counts it reads back: a dense batch (segments > 1024 candidates) between sparse
    ones must still give the oracle's result, synchronously and streamed, and the speculation must resume afterwards."""
    h, w, sizes, B = 540, 960, [640, 960], 2
    sparse = standin.make_batch(h, w, sizes, B, n_faces=20, seed=41, dtype=np.float16)
    dense = standin.make_batch(h, w, sizes, B, n_faces=700, seed=42, dtype=np.float16)

    class Switch(standin.SyntheticHeadModel):
        pass
    model = Switch()
    det = MultiScaleFaceDetector(model, device=DEV, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU, cap=4096)

    def load(batch):
        for s_, hw in enumerate(batch[2]):
            model.set_heads(hw, batch[1][s_], pin=True)

    def check(outs, batch):
        for b in range(B):
            ref, _ = _oracle_detect(batch[1], batch[2], b, (h, w), synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5)
            assert outs[b].shape == ref.shape and np.array_equal(outs[b], ref), (b, outs[b].shape, ref.shape)

    load(sparse)
    check(det.detect_multi_scale_batch(sparse[0]), sparse)
    assert det._speculate_small
    load(dense)
    out = det.detect_multi_scale_batch(dense[0])
    assert det.last_stats["candidates"] > 2 * 1024 * B  # some segment is beyond the single-CTA limit
    check(out, dense)
    assert not det._speculate_small
    check(det.detect_multi_scale_batch(dense[0]), dense)  # no speculation now: one run
    load(sparse)
    check(det.detect_multi_scale_batch(sparse[0]), sparse)
    assert det._speculate_small
    # streamed: the dense batch arrives while the detector speculates
    load(dense)
    for outs in det.detect_stream([dense[0], dense[0]], depth=2):
        check(outs, dense)
