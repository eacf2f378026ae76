"""The UNMODIFIED reference's GPU path timed on the same B200 beside libfdms on the same synthetic workload
("the real bar for NMS").

What runs is the reference's own code (from `baseline/_ref/`, through baseline/ref_runner.py): per image and per scale
`preprocess_image` (cv2 letterbox on the host + H2D + fp16 /255), the `IKeypoint` decode in fp16 ATen elementwise ops,
`non_max_suppression` (python loop, boolean-mask filter, `torchvision.ops.nms` sm_100 kernel, no max_det: the reference
has it commented out, utils/general.py:519,585-586), `scale_coords_api_approach(...).round()`, `.cpu().numpy()`, and the
CPU `weighted_nms`.  The backbone is the stand-in of the bench (raw head maps already resident on the device).
A measurement tool, not part of the product.

Usage: python baseline/ref_gpu_path.py [--workload cfg2|cfg3] [--images 8]
Prints one JSON line (also written to gpurun_out/ref_gpu_path_<workload>.json)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402
from baseline import ref_runner as R  # noqa: E402
from workloads import standin  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--images", type=int, default=8, help="images the reference is timed on (it runs them one by one)")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    strides, anchors = bench.head_geometry(wl)
    dev = torch.device("cuda", 0)
    B, sizes = wl["batch"], wl["sizes"]
    images, heads, shapes = standin.make_batch(wl["h"], wl["w"], sizes, B, wl["faces"], seed=17, radius=wl["radius"],
                                               strides=strides, anchors_px=anchors)
    geoms = [_lib.letterbox_geometry(wl["h"], wl["w"], s, auto=True, stride=int(max(strides))) for s in sizes]
    heads_d = [[torch.from_numpy(l).to(dev) for l in hs] for hs in heads]

    h = R.RefHarness(sizes, wl["conf"], wl["iou"], strides, anchors, device="cuda:0")
    n = min(args.images, B)
    per = [{tuple(int(v) for v in shapes[s]): [l[b] for l in heads_d[s]] for s in range(len(sizes))} for b in range(n)]
    finals = [h.detect(images[b], per[b])[0] for b in range(n)]  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b in range(n):
        h.detect(images[b], per[b])
    torch.cuda.synchronize()
    ms_ref = 1e3 * (time.perf_counter() - t0) / n
    stages = h.stage_times(images[0], per[0])

    img_d = torch.from_numpy(images).to(dev)
    outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
    plan = ops.PostPlan(heads_d, strides, anchors, [g.dst_h for g in geoms], (wl["h"], wl["w"]), wl["conf"], wl["iou"],
                        cap=wl["cap"], out_cap=wl["cap"], iou_mode=ops.IOU_TV_CUDA, with_kpts=bool(wl.get("kpts")))

    def step():
        ops.pyramid(img_d, geoms, out_dtype=torch.float16, layout="chw", flip_rb=True, outs=outs)
        plan.run()
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_fdms = e0.elapsed_time(e1) / 20
    n_out = plan.out.n_out.cpu().numpy()
    out = {"workload": args.workload, "batch": B, "scales": sizes,
           "reference_gpu_ms_per_image": round(ms_ref, 3), "reference_gpu_ms_per_step": round(ms_ref * B, 2),
           "reference_gpu_stages_ms_per_image": stages, "images_timed": n,
           "libfdms_ms_per_step": round(ms_fdms, 4), "speedup": round(ms_ref * B / ms_fdms, 1),
           "final_detections_reference": [int(len(f)) for f in finals],
           "final_detections_libfdms": [int(v) for v in n_out[:n]],
           "note": "reference decodes / filters in fp16 on CUDA, libfdms in fp32 on the same fp16 maps: counts may differ by a few",
           "torch": torch.__version__}
    line = json.dumps(out)
    print(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_gpu_path_%s.json" % args.workload), "w") as f:
        f.write(line + "\n")


if __name__ == "__main__":
    main()
