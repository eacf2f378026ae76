"""Tuning sweep of the pyramid kernel's knobs on one GPU (cfg2 shapes): staged-tile rows (FDMS_PYR_TH), shared-memory
budget (FDMS_PYR_SMEM_KB), direct tiles for integer scales (FDMS_PYR_DIRECT, FDMS_PYR_TH_DIRECT).  Every
configuration must reproduce the first one bit for bit.
Usage: python tools/pyr_sweep.py [B]"""
import itertools
import os
import sys

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
ref_outs = None
configs = [dict(FDMS_PYR_DIRECT=0, FDMS_PYR_TH=8, FDMS_PYR_SMEM_KB=40, FDMS_PYR_TH_DIRECT=8)]
if len(sys.argv) > 2:  # "TH=8,TH_DIRECT=16;TH=4,..." explicit list (names without the FDMS_PYR_ prefix)
    for item in sys.argv[2].split(";"):
        configs.append({"FDMS_PYR_" + kv.split("=")[0]: kv.split("=")[1] for kv in item.split(",")})
else:
    for th, kb, thd in itertools.product((8, 16), (24, 32, 40), (2, 4, 8, 16)):
        configs.append(dict(FDMS_PYR_DIRECT=1, FDMS_PYR_TH=th, FDMS_PYR_SMEM_KB=kb, FDMS_PYR_TH_DIRECT=thd))
DEFAULTS = dict(FDMS_PYR_DIRECT=1, FDMS_PYR_TH=8, FDMS_PYR_SMEM_KB=40, FDMS_PYR_TH_DIRECT=16)
for cfg in configs:
    for k, v in {**DEFAULTS, **cfg}.items():
        os.environ[k] = str(v)
    _lib._tls.ctx = {}
    for o in outs:
        o.zero_()
    for i in range(5):
        ops.pyramid(imgs[i % 2], geoms, outs=outs)
    ops.pyramid(imgs[0], geoms, outs=outs)
    torch.cuda.synchronize()
    if ref_outs is None:
        ref_outs = [o.clone() for o in outs]
    ok = all(torch.equal(a, b) for a, b in zip(outs, ref_outs))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 40
    for i in range(n):
        ops.pyramid(imgs[i % 2], geoms, outs=outs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%s  %.4f ms  %.0f GB/s  %s" % (" ".join("%s=%s" % (k[9:], v) for k, v in cfg.items()), ms, bytes_ / ms / 1e6,
                                          "same" if ok else "DIFFERENT"), flush=True)
