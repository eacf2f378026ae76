"""End-to-end probe: detect_stream (depth 1/2/3) against the synchronous call, repeated, on the cfg2 workload.
Usage: python tools/e2e_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from workloads import standin  # noqa: E402
from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector  # noqa: E402

B, H, W, sizes = 32, 1080, 1920, [640, 960, 1280]
dev = torch.device("cuda", 0)
images, heads, shapes = standin.make_batch(H, W, sizes, B, 60, seed=17, radius=1)
model = standin.SyntheticHeadModel()
for s, hw in enumerate(shapes):
    model.set_heads(hw, heads[s], pin=True)
det = MultiScaleFaceDetector(model, device=dev, img_sizes=sizes, conf_thres=0.5, iou_thres=0.5, cap=int(os.environ.get("PROBE_CAP", "1024")), out_cap=int(os.environ.get("PROBE_CAP", "1024")))
pinned = torch.from_numpy(images).pin_memory()
for _ in range(3):
    det.detect_multi_scale_batch(pinned)
n = 60


def run_sync():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        det.detect_multi_scale_batch(pinned)
    torch.cuda.synchronize()
    return B * n / (time.perf_counter() - t0)


def run_stream(depth):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in det.detect_stream((pinned for _ in range(n)), depth=depth):
        pass
    torch.cuda.synchronize()
    return B * n / (time.perf_counter() - t0)


# raw H2D ceiling: the same bytes, one stream, nothing else
dst_i = torch.empty_like(pinned, device=dev)
heads_p = [t for lv in model._heads.values() for t in lv]
dst_h = [torch.empty_like(t, device=dev) for t in heads_p]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    dst_i.copy_(pinned, non_blocking=True)
    for a, b in zip(dst_h, heads_p):
        a.copy_(b, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
nbytes = pinned.numel() + sum(t.numel() * t.element_size() for t in heads_p)
print("raw H2D: %.1f GB/s -> ceiling %.0f images/s" % (nbytes * n / dt / 1e9, B * n / dt), flush=True)
for rep in range(4):
    print("rep %d: sync %.0f  stream1 %.0f  stream2 %.0f  stream3 %.0f" % (rep, run_sync(), run_stream(1), run_stream(2), run_stream(3)), flush=True)
