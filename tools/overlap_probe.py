"""Which post-processing stage costs the pyramid how much when both streams run?  Pyramid on one stream, ONE stage
(or the whole chain) repeated on the other, cfg2 shapes.  Usage: python tools/overlap_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402
from workloads import standin, synth  # noqa: E402

B, H, W, sizes = 32, 1080, 1920, [640, 960, 1280]
dev = torch.device("cuda", 0)
images, heads, shapes = standin.make_batch(H, W, sizes, B, 60, seed=17, radius=1)
geoms = [_lib.letterbox_geometry(H, W, s, auto=True, stride=32) for s in sizes]
img_d = torch.from_numpy(images).to(dev)
heads_d = [[torch.from_numpy(l).to(dev) for l in hs] for hs in heads]
outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
plan = ops.PostPlan(heads_d, synth.P5_STRIDES, synth.P5_ANCHORS, [g.dst_h for g in geoms], (H, W), 0.5, 0.5, cap=1024,
                    out_cap=1024, iou_mode=ops.IOU_TV_CUDA)
plan.run()
torch.cuda.synchronize()
c = plan.cands
s_pyr, s_post = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
pyr_ctx = _lib.Ctx(0)
in_h = [g.dst_h for g in geoms]
STAGES = {
    "none": lambda: None,
    "decode": lambda: ops.decode_filter(heads_d, synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 1024, out=c, landmarks=False),
    "nms": lambda: ops.nms_segments(c.cand, c.count, c.cap, 0.5, ops.IOU_TV_CUDA, tie=c.row, keep=c.keep, n_keep=c.n_keep),
    "merge": lambda: ops.merge_scales(c, B, 3, in_h, (H, W), 0.5, 1024, out=plan.out),
    "all": lambda: plan.run(),
}
n = 300
for name, fn in STAGES.items():
    for _ in range(5):
        with torch.cuda.stream(s_pyr):
            ops.pyramid(img_d, geoms, outs=outs, ctx=pyr_ctx)
        with torch.cuda.stream(s_post):
            fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    s_pyr.wait_event(e0)
    s_post.wait_event(e0)
    for _ in range(n):
        with torch.cuda.stream(s_pyr):
            ops.pyramid(img_d, geoms, outs=outs, ctx=pyr_ctx)
        with torch.cuda.stream(s_post):
            fn()
    cur.wait_stream(s_pyr)
    cur.wait_stream(s_post)
    e1.record(cur)
    torch.cuda.synchronize()
    print("pyramid || %-7s %.4f ms per step" % (name, e0.elapsed_time(e1) / n), flush=True)
