"""Tuning sweep of the pyramid kernel's tile knobs (FDMS_PYR_SMEM_KB / FDMS_PYR_TH) on one GPU.
Usage: python tools/pyr_sweep.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, W, sizes = 1080, 1920, [640, 960, 1280]
dev = torch.device("cuda", 0)
imgs = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
geoms = [_lib.letterbox_geometry(H, W, s, auto=True, stride=32) for s in sizes]
outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
bytes_ = B * (H * W * 3 + sum(3 * g.dst_h * g.dst_w * 2 for g in geoms))
ref = None
for kb in (56, 40, 32, 24):
    for th in (16, 8, 4):
        os.environ["FDMS_PYR_SMEM_KB"], os.environ["FDMS_PYR_TH"] = str(kb), str(th)
        _lib._tls.ctx = {}
        for i in range(5):
            ops.pyramid(imgs[i % 2], geoms, outs=outs)
        torch.cuda.synchronize()
        chk = [o.view(torch.int16).sum().item() for o in outs]
        if ref is None:
            ref = chk
        assert chk == ref, "results changed with the tile shape"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 40
        for i in range(n):
            ops.pyramid(imgs[i % 2], geoms, outs=outs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("smem_kb=%2d th=%2d  %.4f ms  %.0f GB/s" % (kb, th, ms, bytes_ / ms / 1e6), flush=True)
