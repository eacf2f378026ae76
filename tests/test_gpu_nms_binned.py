"""Large-segment NMS with the kept boxes held in a shared-memory grid (nms_binned_kernel, nms.cu) against the oracle:
keep lists bit-exact (torchvision.ops.nms semantics, utils/general.py:584, multi_scale_face_detector.py:234) on the
distributions that drive each branch of the kernel -- dense crowds (the grid proper), more kept boxes than the grid
holds, boxes covering more cells than one box may, more list entries than the grid holds, degenerate boxes, thresholds
for which disjoint pairs count (plain list sweep), the ncnn +1 pixel convention.  GPU only."""
import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import ops
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODES = [ops.IOU_TV_CPU, ops.IOU_TV_CUDA, ops.IOU_NCNN]


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def crowd(rng, n, n_faces, extent=(1280.0, 720.0), lo=8.0, hi=150.0, jitter=0.05):
    size = np.exp(rng.uniform(np.log(lo), np.log(hi), n_faces))
    ctr = rng.uniform(0.03, 0.97, (n_faces, 2)) * np.array(extent)
    idx = rng.integers(0, n_faces, n)
    wh = np.stack([size[idx] * 0.85, size[idx]], 1) * (1 + rng.normal(0, jitter, (n, 2)))
    c = ctr[idx] + rng.normal(0, jitter, (n, 2)) * wh
    return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)


def run_and_check(boxes_list, scores_list, thr, mode, max_nms=0):
    n_seg = len(boxes_list)
    cap = max(len(b) for b in boxes_list)
    rows = np.zeros((n_seg, cap, 6), np.float32)
    counts = np.zeros((n_seg,), np.int32)
    for g, (b, s) in enumerate(zip(boxes_list, scores_list)):
        rows[g, :len(b), :4] = b
        rows[g, :len(b), 4] = s
        counts[g] = len(b)
    ops.debug_poison_smem(0xFF)
    keep, n_keep = ops.nms_segments(gpu(rows), gpu(counts), cap, thr, mode, max_nms=max_nms)
    torch.cuda.synchronize()
    kept = []
    for g in range(n_seg):
        n = counts[g]
        ref = O.nms(rows[g, :n, :4], rows[g, :n, 4], thr, mode)
        if max_nms:
            order = np.lexsort((np.arange(n), -rows[g, :n, 4].astype(np.float64)))[:max_nms]
            sub = np.sort(order)  # the oracle takes boxes in index order with their scores: feed the cut set
            ref = sub[O.nms(rows[g, sub, :4], rows[g, sub, 4], thr, mode)]
        got = keep[g, :int(n_keep[g])].cpu().numpy()
        assert np.array_equal(got, ref), (g, n, thr, mode, len(got), len(ref))
        kept.append(len(got))
    return kept


@pytest.mark.parametrize("mode", MODES)
def test_dense_crowd_segments(mode):
    rng = np.random.default_rng(10 + mode)
    boxes, scores = [], []
    for g in range(10):
        n = int(rng.integers(1100, 9000))
        boxes.append(crowd(rng, n, int(rng.integers(50, 1500))))
        scores.append(rng.uniform(0.5, 1.0, n).astype(np.float32))
    kept = run_and_check(boxes, scores, 0.5, mode)
    assert max(kept) > 600


@pytest.mark.parametrize("mode", MODES)
def test_more_kept_boxes_than_the_grid_holds(mode):
    """~9 k disjoint small boxes and their duplicates: the keep list passes 4096 boxes, the rest is streamed as a list."""
    rng = np.random.default_rng(20 + mode)
    boxes, scores = [], []
    for g in range(8):
        gx, gy = np.meshgrid(np.arange(110), np.arange(85))
        tl = np.stack([gx.ravel() * 12.0, gy.ravel() * 9.0], 1) + rng.uniform(0, 1.5, (gx.size, 2))
        b = np.concatenate([tl, tl + rng.uniform(6, 8.5, (gx.size, 2))], 1)
        dup = b[rng.integers(0, len(b), 3000)] + rng.normal(0, 0.3, (3000, 4))
        b = np.concatenate([b, dup]).astype(np.float32)
        b = b[rng.permutation(len(b))]
        boxes.append(b)
        scores.append(np.round(rng.uniform(0.3, 1.0, len(b)), 3).astype(np.float32))  # ties
    kept = run_and_check(boxes, scores, 0.45, mode)
    assert min(kept) > 4096


@pytest.mark.parametrize("mode", MODES)
def test_huge_boxes_among_small_ones(mode):
    """A few boxes cover most of the extent (more cells than one box may be linked into): binning stops at the first of
    them that is kept and the list sweep takes over; also every segment of a call ends differently."""
    rng = np.random.default_rng(30 + mode)
    boxes, scores = [], []
    for g in range(9):
        n = 4000 + 300 * g
        b = crowd(rng, n, 700, lo=6.0, hi=40.0)
        nh = 1 + g
        ctr = rng.uniform(300, 500, (nh, 2))
        wh = rng.uniform(500, 1100, (nh, 2))
        b[rng.choice(n, nh, replace=False)] = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1)
        boxes.append(b)
        s = rng.uniform(0.5, 1.0, n).astype(np.float32)
        if g % 2:
            s[:] = np.sort(s)  # the huge boxes arrive late
        scores.append(s)
    run_and_check(boxes, scores, 0.5, mode)


@pytest.mark.parametrize("mode", MODES)
def test_more_list_entries_than_the_grid_holds(mode):
    """Mixed sizes: small boxes set the cell, thousands of larger kept boxes cover 20-60 cells each."""
    rng = np.random.default_rng(40 + mode)
    boxes, scores = [], []
    for g in range(8):
        small = crowd(rng, 5000, 2500, lo=5.0, hi=9.0, jitter=0.02)
        ctr = rng.uniform(0, 1, (1500, 2)) * np.array([1280.0, 720.0])
        wh = rng.uniform(30, 55, (1500, 2))
        big = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
        b = np.concatenate([small, big])
        b = b[rng.permutation(len(b))]
        boxes.append(b)
        scores.append(rng.uniform(0.5, 1.0, len(b)).astype(np.float32))
    run_and_check(boxes, scores, 0.6, mode)


@pytest.mark.parametrize("mode", MODES)
def test_degenerate_boxes_and_integer_corners(mode):
    rng = np.random.default_rng(50 + mode)
    boxes, scores = [], []
    for g in range(8):
        n = 3000
        b = np.round(crowd(rng, n, 400, extent=(400.0, 300.0), lo=4.0, hi=40.0))  # integer corners: touching boxes, IoU on thr
        k = rng.choice(n, 300, replace=False)
        b[k[:100], 2] = b[k[:100], 0]                    # zero width
        b[k[100:200], [3]] = b[k[100:200], [1]] - 5.0    # inverted
        b[k[200:300]] = b[k[200]]                        # exact duplicates
        boxes.append(b.astype(np.float32))
        scores.append(np.round(rng.uniform(0.5, 1.0, n), 2).astype(np.float32))
    for thr in (0.5, 1 / 3, 0.0):
        run_and_check(boxes, scores, thr, mode)


def test_one_box_segments_identical_boxes_and_max_nms():
    rng = np.random.default_rng(60)
    boxes = [np.tile(np.array([[10, 10, 50, 60]], np.float32), (2000, 1)) for _ in range(4)]
    boxes += [crowd(rng, 5000, 900) for _ in range(4)]
    scores = [rng.uniform(0.5, 1.0, len(b)).astype(np.float32) for b in boxes]
    run_and_check(boxes, scores, 0.5, ops.IOU_TV_CUDA)
    run_and_check(boxes, scores, 0.5, ops.IOU_TV_CUDA, max_nms=3000)


def test_grid_and_list_kernels_agree_on_the_bench_distribution(monkeypatch):
    """The A/B switch of DESIGN 4.3: FDMS_NMS_BIN=2 (list kernel), 0 (new kernel without the grid), 1 (grid)."""
    import os
    import subprocess
    import sys
    code = ("import numpy as np, torch, hashlib\n"
            "from face_detection_multi_scale_b200 import ops\n"
            "from tests.test_gpu_nms_binned import crowd\n"
            "rng = np.random.default_rng(7)\n"
            "rows = np.zeros((12, 8000, 6), np.float32)\n"
            "for g in range(12):\n"
            "    rows[g, :, :4] = crowd(rng, 8000, 1300)\n"
            "    rows[g, :, 4] = rng.uniform(0.5, 1, 8000)\n"
            "cnt = np.full((12,), 8000, np.int32)\n"
            "keep, nk = ops.nms_segments(torch.from_numpy(rows).cuda(), torch.from_numpy(cnt).cuda(), 8000, 0.5, ops.IOU_TV_CUDA, max_nms=0)\n"
            "torch.cuda.synchronize()\n"
            "h = hashlib.sha256()\n"
            "for g in range(12): h.update(keep[g, :int(nk[g])].cpu().numpy().tobytes())\n"
            "print(h.hexdigest(), int(nk.sum()))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for v in ("1", "0", "2"):
        env = dict(os.environ, FDMS_NMS_BIN=v, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1] == outs[2], outs


@pytest.mark.parametrize("n", [1500, 3000, 7000])
def test_sort_tie_order_packed_and_wide_keys(n):
    """The register-blocked sort packs the item index under the tie when the ties are small (head row numbers) and
    carries it separately when they are not; exact score ties must resolve by ASCENDING TIE in both cases, whatever
    the order of the rows."""
    rng = np.random.default_rng(70 + n)
    n_seg = 8
    rows = np.zeros((n_seg, n, 6), np.float32)
    for g in range(n_seg):
        rows[g, :, :4] = np.round(crowd(rng, n, 300))
        rows[g, :, 4] = np.round(rng.uniform(0.5, 1.0, n), 2)  # ~50 distinct scores: ties everywhere
    counts = np.full((n_seg,), n, np.int32)
    idx = np.arange(n, dtype=np.int64)
    for name, tie_of in (("small ascending", idx * 3 + 5), ("small descending", (n - 1 - idx) * 7),
                         ("wide ascending", idx * 1000 + (1 << 21)), ("wide descending", (n - 1 - idx) * 4099 + (1 << 20))):
        tie = np.tile(tie_of.astype(np.int32), (n_seg, 1))
        keep, n_keep = ops.nms_segments(gpu(rows), gpu(counts), n, 0.5, ops.IOU_TV_CUDA, max_nms=0, tie=gpu(tie))
        torch.cuda.synchronize()
        for g in range(n_seg):
            if "ascending" in name:
                ref = O.nms(rows[g, :, :4], rows[g, :, 4], 0.5, ops.IOU_TV_CUDA)
            else:  # the oracle breaks ties by position: feed it the rows in reverse
                ref = n - 1 - O.nms(rows[g, ::-1, :4], rows[g, ::-1, 4], 0.5, ops.IOU_TV_CUDA)
            got = keep[g, :int(n_keep[g])].cpu().numpy()
            assert np.array_equal(got, ref), (name, g, n)
