#!/usr/bin/env python
"""Diagnostic for tests/test_gpu_parity.py::test_postprocess_fuzz_vs_oracle: re-runs one seeded case and prints where
the merged detections differ from the oracle's (rows, columns, values, the unrounded boxes of the oracle).

    python tests/fuzz_postprocess_diag.py SEED ITERATION
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from face_detection_multi_scale_b200 import ops  # noqa: E402
from workloads import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (test infrastructure)


def main():
    seed, want = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(500 + seed)
    for it in range(want + 1):
        p6 = bool(rng.integers(0, 2))
        strides, anchors = (synth.P6_STRIDES, synth.P6_ANCHORS) if p6 else (synth.P5_STRIDES, synth.P5_ANCHORS)
        smax = strides[-1]
        n_scales, bs = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        dtype = np.float16 if rng.integers(0, 2) else np.float32
        shapes = [(int(rng.integers(1, 6)) * smax, int(rng.integers(1, 7)) * smax) for _ in range(n_scales)]
        conf = float(rng.choice([0.05, 0.25, 0.5, 0.7]))
        iou = float(rng.choice([0.3, 0.45, 0.5]))
        mode = ops.IOU_TV_CPU if rng.integers(0, 2) else ops.IOU_TV_CUDA
        h0, w0 = int(rng.integers(40, 2000)), int(rng.integers(40, 2000))
        bias = float(rng.choice([-3.0, -1.5, 0.0]))
        sl = []
        for ih, iw in shapes:
            lv = []
            for s in strides:
                a = (rng.standard_normal((bs, 63, ih // s, iw // s)) * 1.5).astype(np.float32)
                a[:, 4::21] += bias
                lv.append(a.astype(dtype))
            sl.append(lv)
    print("case", seed, want, "p6", p6, shapes, "bs", bs, dtype, "conf", conf, "iou", iou, "mode", mode, "orig", (h0, w0))
    rows_max = max(sum(3 * (ih // s) * (iw // s) for s in strides) for ih, iw in shapes)
    dev = "cuda:0"
    plan = ops.PostPlan([[torch.from_numpy(l).to(dev) for l in lv] for lv in sl], strides, anchors,
                        [s[0] for s in shapes], (h0, w0), conf, iou, cap=rows_max, out_cap=rows_max * n_scales,
                        iou_mode=mode)
    out = plan.run()
    torch.cuda.synchronize()
    for b in range(bs):
        heads_b = [[l[b:b + 1] for l in lv] for lv in sl]
        ref, per_scale = O.detect_from_heads(heads_b, shapes, (h0, w0, 3), strides, anchors, conf, iou, mode)
        n = int(out.n_out[b])
        got = out.dets[b, :n].cpu().numpy().astype(np.float64)
        print("image", b, "got", got.shape, "ref", ref.shape)
        if got.shape != ref.shape:
            continue
        bad = np.argwhere(got != ref)
        print("differing entries:", len(bad), "rows:", sorted(set(bad[:, 0].tolist()))[:20])
        for r in sorted(set(bad[:, 0].tolist()))[:8]:
            print(" row", r, "\n  got", got[r].tolist(), "\n  ref", ref[r].tolist())
        # same multiset of rows?
        gs = got[np.lexsort(got.T[::-1])]
        rs = ref[np.lexsort(ref.T[::-1])]
        print("same rows up to order:", bool(np.array_equal(gs, rs)))
        # GPU candidates against the oracle's decode for the first scale
        pred = O.decode_head(heads_b[0], strides, anchors)
        rc, rr = O.filter_xyxy(pred[0], conf)
        c = plan.cands
        seg = b * n_scales
        k = int(c.count[seg])
        rows = c.row[seg, :k].cpu().numpy()
        cand = c.cand[seg, :k].cpu().numpy()
        o = np.argsort(rows, kind="stable")
        d = np.abs(cand[o][:, :6] - rc[:, :6])
        print("scale 0 candidates:", k, "max |diff| cols 0..5:", d.max(0).tolist(), "differing values:", int((d > 0).sum()))


if __name__ == "__main__":
    main()
