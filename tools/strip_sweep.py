"""A/B and tuning sweep of the pyramid strip kernel on one GPU: FDMS_PYR_STRIP=0 (tile kernel) against the strip kernel
over ring size, stage size and consumer warps.  Every configuration must reproduce the tile kernel bit for bit.
Usage: python tools/strip_sweep.py [cfg2|cfg4|cfg5] [B] ["STRIP=1,STRIP_SMEM_KB=144;..."]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402

SHAPES = {"cfg2": (1080, 1920, [640, 960, 1280], 32, 32), "cfg4": (2160, 3840, [640, 960, 1280, 1536], 64, 16),
          "cfg5": (1080, 1920, [640], 32, 64)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
H, W, sizes, stride, B = SHAPES[name]
if len(sys.argv) > 2 and int(sys.argv[2]) > 0:
    B = int(sys.argv[2])
dev = torch.device("cuda", 0)
imgs = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
geoms = [_lib.letterbox_geometry(H, W, s, auto=True, stride=stride) for s in sizes]
outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
bytes_ = B * (H * W * 3 + sum(3 * g.dst_h * g.dst_w * 2 for g in geoms))
configs = [dict(STRIP=0)]
if len(sys.argv) > 3:
    for item in sys.argv[3].split(";"):
        configs.append({kv.split("=")[0]: kv.split("=")[1] for kv in item.split(",")})
else:
    for smem in (72, 108, 144, 180, 208):
        for warps in (8, 12, 16):
            configs.append(dict(STRIP=1, STRIP_SMEM_KB=smem, STRIP_WARPS=warps))
    configs.append(dict(STRIP=1, STRIP_STAGE_KB=18, STRIP_SMEM_KB=144))
ref = None
for cfg in configs:
    for k in ("STRIP", "STRIP_SMEM_KB", "STRIP_STAGE_KB", "STRIP_WARPS"):
        os.environ.pop("FDMS_PYR_" + k, None)
    for k, v in cfg.items():
        os.environ["FDMS_PYR_" + k] = str(v)
    _lib._tls.ctx = {}
    for o in outs:
        o.fill_(float("nan"))
    for i in range(5):
        ops.pyramid(imgs[i % 2], geoms, outs=outs)
    ops.pyramid(imgs[0], geoms, outs=outs)
    torch.cuda.synchronize()
    if ref is None:
        ref = [o.clone() for o in outs]
    ok = all(torch.equal(a.view(torch.int16), b.view(torch.int16)) for a, b in zip(outs, ref))
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 40
        for i in range(n):
            ops.pyramid(imgs[i % 2], geoms, outs=outs)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    print("%-50s %.4f ms  %.0f GB/s  %s" % (" ".join("%s=%s" % kv for kv in cfg.items()), best, bytes_ / best / 1e6,
                                            "same" if ok else "DIFFERENT"), flush=True)
