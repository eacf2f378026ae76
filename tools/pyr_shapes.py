"""Pyramid kernel over a few source shapes / scale sets (algorithmic GB/s of each).  Usage: python tools/pyr_shapes.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
CASES = [("cfg2 1080p", 32, 1080, 1920, [640, 960, 1280], False),
         ("cfg2 1080p api", 32, 1080, 1920, [640, 960, 1280], True),
         ("cfg4 4K P6", 16, 2160, 3840, [640, 960, 1280, 1536], False),
         ("cfg4 4K P6 api", 16, 2160, 3840, [640, 960, 1280, 1536], True),
         ("cfg5 1080p single", 64, 1080, 1920, [640], False),
         ("720p", 32, 720, 1280, [640, 960, 1280], False),
         ("1080p odd sizes", 32, 1080, 1920, [736, 1088], False),
         ("portrait 1080p", 32, 1920, 1080, [640, 960, 1280], False),
         ("portrait 1080p odd sizes", 32, 1920, 1080, [736, 1088], False),
         ("portrait 1080p api", 32, 1920, 1080, [640, 960, 1280], True),
         ("1366x768", 32, 768, 1366, [640, 960, 1280], False),
         ("prod [640, 3840] api", 8, 1080, 1920, [640, 3840], True)]
for name, B, H, W, sizes, api in CASES:
    imgs = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    S = max(H, W)
    geoms = [_lib.letterbox_geometry(S, S, s, auto=False, stride=64) if api else _lib.letterbox_geometry(H, W, s, auto=True, stride=64)
             for s in sizes]
    outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
    bytes_ = B * (H * W * 3 + sum(3 * g.dst_h * g.dst_w * 2 for g in geoms))
    kw = dict(virt_hw=(S, S)) if api else {}
    for i in range(3):
        ops.pyramid(imgs[i % 2], geoms, outs=outs, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        ops.pyramid(imgs[i % 2], geoms, outs=outs, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%-24s B=%2d -> %-24s %8.4f ms  %6.0f GB/s  (%.0f MB)" % (name, B, [(g.dst_h, g.dst_w) for g in geoms], ms,
                                                                    bytes_ / ms / 1e6, bytes_ / 1e6), flush=True)
    del imgs, outs
    torch.cuda.empty_cache()
