#!/usr/bin/env python
"""Backbone time on its own line and the end-to-end rate that includes it.

Run by bench.py (rank 0, N=1) as a subprocess after its own measurements, or by hand on the GPU box:

    python baseline/backbone_bench.py [--cfg cfg/yolov7s-face.yaml] [--batch 32] [--sizes 640,960,1280]

A REAL reference network -- random-init `models.yolo.Model(cfg)` from the reference tree (`baseline/_ref/` on the GPU
box), fp16 as the reference runs it on CUDA (msfd:49) -- stands between the two halves of the hot path:
  backbone_ms_per_step         the PyTorch forward of the unchanged network, all scales, batch B (CUDA events)
  libfdms                      MultiScaleFaceDetector.detect_multi_scale_batch(pinned host batch): H2D + pyramid + backbone + decode /
                               NMS / merge + D2H, and the same with the forward time subtracted (= the hot path's share);
                               detect_stream over the same batches (copies of neighbouring batches overlap the backbone)
  reference_gpu                the UNMODIFIED reference on the same GPU and the same network: detect_multi_scale per image
                               (msfd:242-288; image served from memory, no file decode), batch 1, sequential scales
Prints one JSON line.  Plain random-init heads give no candidates at conf 0.5: the network's BatchNorm
statistics are set from random batches (baseline/ref_runner.py:random_init_model), after which about 1-2 % of the rows pass
the reference's default conf_thres 0.5 -- the candidate density of cfg2.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg/yolov7s-face.yaml")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--sizes", default="640,960,1280")
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--conf", type=float, default=0.5)
    args = ap.parse_args()
    from baseline import refshim
    if not refshim.install():
        print(json.dumps({"unavailable": "no reference tree (baseline/_ref not installed)"}))
        return 0
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"unavailable": "no CUDA device"}))
        return 0
    warnings.filterwarnings("ignore", message="torch.meshgrid")
    with contextlib.redirect_stdout(io.StringIO()):
        import multi_scale_face_detector as msfd
    from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector as Mine
    from baseline.ref_runner import _Cv2Mem
    import cv2

    sizes = [int(v) for v in args.sizes.split(",")]
    B, dev = args.batch, torch.device("cuda", 0)
    from baseline.ref_runner import random_init_model
    net = random_init_model(args.cfg, seed=0)
    ckpt = os.path.join(tempfile.mkdtemp(prefix="fdms_bb_"), "random_init.pt")
    torch.save({"model": net}, ckpt)
    rng = np.random.default_rng(3)
    images = rng.integers(0, 256, (B, args.h, args.w, 3), dtype=np.uint8)
    pinned = torch.from_numpy(images).pin_memory()

    # the reference's own detector (its constructor, its checkpoint loader), half precision on CUDA
    with contextlib.redirect_stdout(io.StringIO()):
        ref = msfd.MultiScaleFaceDetector(ckpt, "", sizes, 0.5, 0.5)
    store = {}
    msfd.cv2 = _Cv2Mem(cv2, store)
    for b in range(B):
        store["mem:%d" % b] = images[b]
    conf = args.conf
    ref.conf_thres = conf

    mine = Mine(ckpt, "", sizes, conf, 0.5, half=True, cap=16384)
    for _ in range(2):
        out = mine.detect_multi_scale_batch(pinned)
    stats = dict(mine.last_stats)

    def clock(fn, n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    t_mine = clock(lambda: mine.detect_multi_scale_batch(pinned), args.steps)
    # the same through the pipelined entry: the H2D copy of batch i+1 and the D2H of batch i-1 overlap the backbone of batch i
    pinned2 = torch.from_numpy(rng.integers(0, 256, images.shape, dtype=np.uint8)).pin_memory()
    n_stream = max(4, args.steps)

    def stream_run():
        for _ in mine.detect_stream((pinned if i % 2 == 0 else pinned2 for i in range(n_stream)), depth=2):
            pass
    stream_run()
    t_stream = clock(stream_run, 1) / n_stream
    # backbone alone: the forward of the unchanged network on the pyramid's tensors, all scales (CUDA events)
    xs, _ = mine._pyramid(pinned.to(dev))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_scale = {}
    with torch.no_grad():
        for s, x in zip(sizes, xs):
            mine._raw_heads(x)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                mine._raw_heads(x)
            e1.record()
            torch.cuda.synchronize()
            per_scale[str(s)] = round(e0.elapsed_time(e1) / args.steps, 3)
    bb_ms = sum(per_scale.values())
    del xs
    # the unmodified reference on the same GPU: per image, per scale
    n_ref = min(B, 8)

    def ref_run():
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            for b in range(n_ref):
                ref.detect_multi_scale("mem:%d" % b)
    ref_run()
    t_ref = clock(ref_run, 2) / n_ref
    print(json.dumps({
        "model": args.cfg, "weights": "random init (seed 0) + BatchNorm statistics from random batches, fp16", "batch": B, "scales": sizes,
        "source_hw": [args.h, args.w], "conf_thres": conf,
        "candidates_per_step": stats.get("candidates"), "final_detections_per_step": stats.get("final"),
        "backbone_ms_per_step": round(bb_ms, 2), "backbone_ms_per_scale": per_scale,
        "backbone_images_per_s": round(B / (bb_ms * 1e-3), 1),
        "libfdms": {"e2e_with_backbone_images_per_s": round(B / t_mine, 1), "ms_per_step": round(1e3 * t_mine, 2),
                    "hot_path_ms_per_step": round(1e3 * t_mine - bb_ms, 2),
                    "detect_stream_images_per_s": round(B / t_stream, 1), "detect_stream_ms_per_step": round(1e3 * t_stream, 2),
                    "detect_stream_api": "MultiScaleFaceDetector.detect_stream(pinned host batches, depth=2): copies overlapped with the "
                                         "backbone of the neighbouring batches",
                    "api": "MultiScaleFaceDetector.detect_multi_scale_batch(pinned host batch): H2D + pyramid + backbone + post + D2H"},
        "reference_gpu": {"e2e_with_backbone_images_per_s": round(1.0 / t_ref, 2), "ms_per_image": round(1e3 * t_ref, 2),
                          "images_timed": n_ref,
                          "api": "unmodified reference MultiScaleFaceDetector.detect_multi_scale on cuda (fp16), image from memory"},
        "speedup_e2e_with_backbone": round((B / t_mine) * t_ref, 2)}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
