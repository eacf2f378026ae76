"""Run the pyramid kernel a few times on the cfg2 shapes (profiling target).  Usage: python tools/pyr_once.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B, H, W, sizes = 32, 1080, 1920, [640, 960, 1280]
dev = torch.device("cuda", 0)
imgs = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
geoms = [_lib.letterbox_geometry(H, W, s, auto=True, stride=32) for s in sizes]
outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
for i in range(n):
    ops.pyramid(imgs[i % 2], geoms, outs=outs)
torch.cuda.synchronize()
print("ok")
