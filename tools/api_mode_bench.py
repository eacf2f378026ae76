"""Pyramid kernel in the reference's production ("api") preprocessing mode: zero-padded virtual square.
Usage: python tools/api_mode_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402

B, H, W, sizes = 32, 1080, 1920, [640, 960, 1280]
dev = torch.device("cuda", 0)
imgs = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
S = max(H, W)
geoms = [_lib.letterbox_geometry(S, S, s, auto=False, stride=32) for s in sizes]
outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
bytes_ = B * (H * W * 3 + sum(3 * g.dst_h * g.dst_w * 2 for g in geoms))
for i in range(5):
    ops.pyramid(imgs[i % 2], geoms, outs=outs, virt_hw=(S, S))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 40
for i in range(n):
    ops.pyramid(imgs[i % 2], geoms, outs=outs, virt_hw=(S, S))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("api mode, B=%d 1080p -> %s squares: %.4f ms  %.0f GB/s algorithmic (%.1f MB)" % (B, sizes, ms, bytes_ / ms / 1e6, bytes_ / 1e6))
