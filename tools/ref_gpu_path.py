"""The reference's own GPU post-processing recipe, restated in stock PyTorch + torchvision, timed on the same B200
beside libfdms on the same synthetic workload ("the real bar for NMS").

What the reference runs on a CUDA device for every image and every scale (multi_scale_face_detector.py:99-131):
    head decode in torch elementwise ops   (models/yolo.py:267-308)
    non_max_suppression                    (utils/general.py:506-600): python loop over images, boolean-mask
                                           filter, xywh2xyxy, torchvision.ops.nms (sm_100 kernel + host sync)
    scale_coords                           (utils/general.py:374-392)
    .cpu().numpy()                         (multi_scale_face_detector.py:128)
The letterbox stays on the CPU in the reference (cv2) and is not part of this comparison; the cross-scale merge is
CPU numpy there and is not timed here either.  This script is a measurement tool, not part of the product, and
touches neither `oracle/` nor `/root/reference`.

Usage: python tools/ref_gpu_path.py [--workload cfg2|cfg3] [--reps 5] [--half]
Prints one JSON line (also written to gpurun_out/ref_gpu_path_<workload>.json)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from face_detection_multi_scale_b200 import _lib, ops  # noqa: E402
from workloads import standin, synth  # noqa: E402

WORKLOADS = {
    "cfg2": dict(batch=32, h=1080, w=1920, sizes=[640, 960, 1280], faces=60, conf=0.5, iou=0.5, cap=1024, radius=1),
    "cfg3": dict(batch=32, h=1080, w=1920, sizes=[640, 960, 1280], faces=1300, conf=0.02, iou=0.5, cap=16384, radius=1),
}


def torch_decode(levels, strides, anchors_px):
    """Stock-PyTorch restatement of the IKeypoint inference branch: list of (bs, 63, ny, nx) -> (bs, rows, 21)."""
    z = []
    for x, s, anc in zip(levels, strides, anchors_px):
        bs, _, ny, nx = x.shape
        x = x.view(bs, 3, 21, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        gy, gx = torch.meshgrid(torch.arange(ny, device=x.device), torch.arange(nx, device=x.device), indexing="ij")
        grid = torch.stack((gx, gy), 2).view(1, 1, ny, nx, 2).to(x.dtype)
        ag = torch.as_tensor(np.asarray(anc, np.float32), device=x.device).to(x.dtype).view(1, 3, 1, 1, 2)
        det, kpt = x[..., :6], x[..., 6:].clone()
        y = det.sigmoid()
        xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * s
        wh = (y[..., 2:4] * 2.0) ** 2 * ag
        kpt[..., 0::3] = (kpt[..., 0::3] * 2.0 - 0.5 + grid[..., 0:1].repeat(1, 1, 1, 1, 5)) * s
        kpt[..., 1::3] = (kpt[..., 1::3] * 2.0 - 0.5 + grid[..., 1:2].repeat(1, 1, 1, 1, 5)) * s
        kpt[..., 2::3] = kpt[..., 2::3].sigmoid()
        z.append(torch.cat((xy, wh, y[..., 4:], kpt), -1).view(bs, -1, 21))
    return torch.cat(z, 1)


def torch_nms_loop(pred, conf_thres, iou_thres, max_nms=30000, max_det=300):
    """Stock restatement of non_max_suppression for nc=1 with landmarks: one python iteration per image."""
    out = []
    xc = pred[..., 4] > conf_thres
    for xi in range(pred.shape[0]):
        x = pred[xi][xc[xi]]
        if not x.shape[0]:
            out.append(x.new_zeros((0, 21)))
            continue
        x = x.clone()
        x[:, 5:6] *= x[:, 4:5]
        box = x[:, :4].clone()
        box[:, 0] = x[:, 0] - x[:, 2] / 2
        box[:, 1] = x[:, 1] - x[:, 3] / 2
        box[:, 2] = x[:, 0] + x[:, 2] / 2
        box[:, 3] = x[:, 1] + x[:, 3] / 2
        conf, j = x[:, 5:6].max(1, keepdim=True)
        x = torch.cat((box, conf, j.to(x.dtype), x[:, 6:]), 1)[conf.view(-1) > conf_thres]
        n = x.shape[0]
        if not n:
            out.append(x.new_zeros((0, 21)))
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * 4096
        keep = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)
        if keep.shape[0] > max_det:
            keep = keep[:max_det]
        out.append(x[keep])
    return out


def torch_scale_coords(in_hw, coords, img0_hw):
    gain = min(in_hw[0] / img0_hw[0], in_hw[1] / img0_hw[1])
    pad = (in_hw[1] - img0_hw[1] * gain) / 2, (in_hw[0] - img0_hw[0] * gain) / 2
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    coords[:, 0].clamp_(0, img0_hw[1])
    coords[:, 1].clamp_(0, img0_hw[0])
    coords[:, 2].clamp_(0, img0_hw[1])
    coords[:, 3].clamp_(0, img0_hw[0])
    return coords


def wall(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--half", action="store_true", help="decode in fp16 like the reference's CUDA path (msfd:48,101)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    B, sizes = wl["batch"], wl["sizes"]
    images, heads, shapes = standin.make_batch(wl["h"], wl["w"], sizes, B, wl["faces"], seed=17, radius=wl["radius"])
    geoms = [_lib.letterbox_geometry(wl["h"], wl["w"], s, auto=True, stride=32) for s in sizes]
    dt = torch.float16 if args.half else torch.float32
    heads_ref = [[torch.from_numpy(l).to(dev).to(dt) for l in hs] for hs in heads]
    heads_d = [[torch.from_numpy(l).to(dev) for l in hs] for hs in heads]

    # ---- the reference recipe: per scale decode, per image NMS + rescale + D2H ----
    def ref_decode():
        return [torch_decode(hs, synth.P5_STRIDES, synth.P5_ANCHORS) for hs in heads_ref]

    ms_dec, preds = wall(ref_decode, args.reps)
    preds = [p.float() for p in preds]

    def ref_nms():
        return [torch_nms_loop(p, wl["conf"], wl["iou"]) for p in preds]

    ms_nms, kept = wall(ref_nms, args.reps)

    def ref_tail():
        res = []
        for s, g in enumerate(geoms):
            for d in kept[s]:
                d = d.clone()
                torch_scale_coords((g.dst_h, g.dst_w), d[:, :4], (wl["h"], wl["w"]))
                res.append(d[:, :6].cpu().numpy())
        return res

    ms_tail, _ = wall(ref_tail, args.reps)
    n_cand = int(sum(int((p[..., 4] > wl["conf"]).sum()) for p in preds))
    n_keep_ref = int(sum(d.shape[0] for ks in kept for d in ks))

    # ---- libfdms on the same head maps ----
    plan = ops.PostPlan(heads_d, synth.P5_STRIDES, synth.P5_ANCHORS, [g.dst_h for g in geoms], (wl["h"], wl["w"]),
                        wl["conf"], wl["iou"], cap=wl["cap"], out_cap=wl["cap"], iou_mode=ops.IOU_TV_CUDA)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms_fdms = e0.elapsed_time(e1) / n
    c = plan.cands
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_dec = t_nms = 0.0
    for _ in range(n):
        ev[0].record()
        ops.decode_filter(heads_d, synth.P5_STRIDES, synth.P5_ANCHORS, wl["conf"], wl["cap"], out=c)
        ev[1].record()
        ops.nms_segments(c.cand, c.count, c.cap, wl["iou"], ops.IOU_TV_CUDA, tie=c.row, keep=c.keep, n_keep=c.n_keep)
        ev[2].record()
        torch.cuda.synchronize()
        t_dec += ev[0].elapsed_time(ev[1]) / n
        t_nms += ev[1].elapsed_time(ev[2]) / n
    n_keep = int(c.n_keep.sum().item())
    n_cand_fdms = int(c.count.sum().item())

    out = {
        "workload": args.workload, "batch": B, "scales": sizes, "ref_dtype": "f16" if args.half else "f32",
        "candidates": n_cand, "candidates_fdms": n_cand_fdms, "kept_ref": n_keep_ref, "kept_fdms": n_keep,
        "reference_gpu_recipe_ms": {"decode_torch": ms_dec, "nms_loop_torchvision": ms_nms, "scale_coords_d2h": ms_tail,
                                    "total": ms_dec + ms_nms + ms_tail},
        "libfdms_ms": {"decode_filter": t_dec, "nms": t_nms, "postprocess_total_incl_merge": ms_fdms},
        "speedup_post": (ms_dec + ms_nms + ms_tail) / ms_fdms,
        "speedup_nms": ms_nms / t_nms,
        "torch": torch.__version__, "torchvision": torchvision.__version__,
    }
    line = json.dumps(out)
    print(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_gpu_path_%s%s.json" % (args.workload, "_half" if args.half else "")), "w") as f:
        f.write(line + "\n")


if __name__ == "__main__":
    main()
