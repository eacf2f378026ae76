#!/usr/bin/env python
"""Benchmark of the multi-scale face-detection pre/post-processing hot path on B200 (libfdms).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--configs all|none] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input: the fused pyramid kernel over B source
images (all scales, one launch) followed by the post-processing of the B x scales raw head tensors (decode+filter,
batched NMS, rescale, cross-scale weighted NMS).  The PyTorch backbone between the two halves is out of scope
(BASELINE.json north_star: it stays unchanged) and is NOT in the timed region of the headline; its outputs are synthetic
raw head maps with planted faces (random-init weights yield no candidates).  The backbone is measured on its
own line (`backbone`, a real random-init `models.yolo.Model` of the reference, N=1) together with the end-to-end rate
that includes it.

Prints ONE JSON line (rank 0):
  value          images/s with all inputs resident in HBM: the MEDIAN over `repeats` timed regions of EXACTLY K steps each
                 (barrier + synchronize on both sides, CUDA events, max over ranks); every region is listed
  e2e            the same metric through the public API (MultiScaleFaceDetector.detect_stream) with pinned HOST inputs:
                 H2D of images and head maps and D2H of the detections inside the timed region; `h2d_ceiling` = the
                 host->device rate of this box measured in the same run with all ranks copying at once
  roofline       algorithmic bytes of a pyramid launch / its duration (CUDA events on the kernel's own stream inside the
                 timed region; `isolated` = the kernel alone) against MEASURED_PEAKS.json
  configs        short runs of the other BASELINE.json configs (cfg1, cfg3, cfg4, cfg5): value, stages, candidates/s, roofline
  backbone       (N=1) a real random-init reference network between the two halves: backbone ms, end to end with it
  ingest         (N=1) from-files decode: device JPEG ingest against the host decoder (tools/ingest_bench.py)
  cpu_baseline   (N=1) the UNMODIFIED reference on the host cores of this box (baseline/ref_runner.py, kind "reference";
                 kind "port" = the same cv2 / torch / torchvision calls when no reference tree travelled), both CPU modes
`--impl reference` prints the reference arm alone, in the same format.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "cfg1": dict(desc="cfg1: yolov7-tiny-face head, single 640x640 synthetic image, single scale 640",
                 h=640, w=640, sizes=[640], batch=1, faces=20, conf=0.5, iou=0.5, radius=1, cap=4096, head="p5"),
    # configs[1]: the configuration the metric is quoted on
    "cfg2": dict(desc="cfg2: yolov7s-face head, 3-scale pyramid (640/960/1280), batch 32 synthetic 1920x1080 images",
                 h=1080, w=1920, sizes=[640, 960, 1280], batch=32, faces=60, conf=0.5, iou=0.5, radius=1, cap=4096, head="p5"),
    # configs[2]: dense crowd, ~20k pre-NMS candidates per image, landmarks on
    "cfg3": dict(desc="cfg3: dense-crowd stress, ~20k pre-NMS candidates/image, landmarks on, 3-scale pyramid, batch 32 1920x1080",
                 h=1080, w=1920, sizes=[640, 960, 1280], batch=32, faces=1300, conf=0.5, iou=0.5, radius=1, cap=16384,
                 kpts=True, head="p5"),
    # configs[3]: P6 head, 4 scales up to 1536, 4K sources
    "cfg4": dict(desc="cfg4: yolov7-w6 P6 head, 4-scale pyramid (640/960/1280/1536), batch 16 synthetic 3840x2160 images",
                 h=2160, w=3840, sizes=[640, 960, 1280, 1536], batch=16, faces=120, conf=0.5, iou=0.5, radius=1, cap=4096,
                 head="p6"),
    # configs[4]: single-scale high throughput (per-GPU shard of the batch-512 job: 64 images per GPU at 8 GPUs)
    "cfg5": dict(desc="cfg5: yolov7-lite-t head, single scale 640, batch 512 image-sharded over 8 GPUs (64 per GPU) synthetic 1920x1080",
                 h=1080, w=1920, sizes=[640], batch=64, faces=40, conf=0.5, iou=0.5, radius=1, cap=4096, head="p5"),
}
METRIC = "images/sec (3-scale pyramid, device-timed)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="fdms", choices=["fdms", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--configs", default="all", help="'all', 'none' or a comma list of the other configs to add to the line")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--cap", type=int, default=0, help="override the candidate capacity per (image, scale)")
    ap.add_argument("--repeats", type=int, default=0, help="timed regions (0 = 15 when a region is shorter than 50 ms, else 3)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of all-cores CPU baseline inside the fdms arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-backbone", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time plain launches instead of CUDA-graph replays")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: main workload only, no CPU baseline, no e2e, no clocks")
    return ap.parse_args()


def head_geometry(wl):
    from workloads import synth
    return (synth.P6_STRIDES, synth.P6_ANCHORS) if wl["head"] == "p6" else (synth.P5_STRIDES, synth.P5_ANCHORS)


def make_config(wl, world):
    """The `config` object: identical in the fdms arm and in the reference arm (static description of the workload)."""
    return {"workload": wl["desc"], "batch_per_gpu": wl["batch"], "source_hw": [wl["h"], wl["w"]], "scales": wl["sizes"],
            "head": wl["head"], "conf_thres": wl["conf"], "iou_thres": wl["iou"], "planted_faces_per_image": wl["faces"],
            "landmarks": bool(wl.get("kpts")),
            "backbone": "excluded (unchanged PyTorch, out of scope); raw head maps are synthetic fp16",
            "l2": "inputs larger than L2 and two input sets rotated",
            "parallelism": "image-sharded over %d GPUs, no collective" % world if world > 1 else "single GPU"}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------- CPU reference
def reference_block(wl, steps, warmup, budget_s, as_written_images=8):
    """The reference's CPU path on this box's host cores over the workload `wl` (baseline/ref_runner.py).
    all_cores: one spawned single-thread worker per host thread, 4 images each per step; `steps` steps at most, stopped
    after `budget_s` seconds.  as_written: one process, sequential.  Also the scalar C oracle for comparison."""
    from baseline import ref_runner as R
    from workloads import standin
    strides, anchors = head_geometry(wl)
    kind = "reference" if R.available() else "port"
    thr = host_threads()
    out = {"kind": kind, "cpu_model": cpu_model(), "host_threads": thr, "os_cpu_count": os.cpu_count()}
    try:
        import torch
        out["torch_threads"] = torch.get_num_threads()
    except Exception:
        pass
    images, heads, shapes = standin.make_batch(wl["h"], wl["w"], wl["sizes"], min(as_written_images, 4), wl["faces"], seed=17,
                                               radius=wl["radius"], strides=strides, anchors_px=anchors)
    # (i) as written: one process, batch 1, sequential scales, default torch threads, OpenCV single-threaded
    if kind == "reference":
        h = R.RefHarness(wl["sizes"], wl["conf"], wl["iou"], strides, anchors)
        R.run_as_written(wl, images, heads, shapes, strides, anchors, 1, harness=h)
        secs, _ = R.run_as_written(wl, images, heads, shapes, strides, anchors, as_written_images, harness=h)
        out["stages_ms_per_image"] = h.stage_times(images[0], R.maps_of(heads, shapes, 0))
    else:
        R.run_port(wl, images, heads, shapes, strides, anchors, 1)
        secs, _ = R.run_port(wl, images, heads, shapes, strides, anchors, as_written_images)
    out["as_written"] = {"value": round(as_written_images / secs, 2), "unit": "images/s", "ms_per_image": round(1e3 * secs / as_written_images, 2),
                         "cores": out.get("torch_threads", 1),
                         "sample": "%d images, one process, batch 1, sequential scales (multi_scale_face_detector.py:242-288)" % as_written_images}
    # the scalar C oracle on the same image (what round 1 reported as the baseline), one thread
    try:
        from oracle import oracle as O
        O.lib()
        t0 = time.perf_counter()
        for size in wl["sizes"]:
            lb, _, _ = O.letterbox(images[0], size, auto=True, stride=int(max(strides)))
            O.pack_chw(lb, flip=True, dtype=np.float16)
        O.detect_from_heads([[l[0:1] for l in hs] for hs in heads], shapes, images[0].shape, strides, anchors, wl["conf"], wl["iou"], O.TV_CPU)
        out["c_oracle_ms_per_image"] = round(1e3 * (time.perf_counter() - t0), 2)
    except Exception as e:  # the oracle is only a comparison point here
        out["c_oracle_ms_per_image"] = "unavailable: %r" % (e,)
    # (ii) all cores: spawned single-thread workers, image-sharded
    per_worker = 4
    pool = R.RefPool(wl, strides, anchors, thr, per_worker, use_port=(kind != "reference"))
    try:
        for _ in range(max(1, warmup)):
            pool.step()
        dt, done = 0.0, 0
        while done < steps and (dt < budget_s or done == 0):
            s, _ = pool.step()
            dt += s
            done += 1
    finally:
        pool.close()
    out.update({"value": round(pool.images_per_step * done / dt, 2), "unit": "images/s", "cores": thr, "steps_run": done,
                "images_per_step": pool.images_per_step, "ms_per_step": round(1e3 * dt / done, 2),
                "sample": "%d steps x %d images (%d single-thread worker processes x %d images, spawn start method, image-sharded like "
                          "yolov7_face_multi_scale_dataframe_predict.py:1115,1147), all stages of multi_scale_face_detector.py:242-288 "
                          "around stand-in head maps, %.1f s" % (done, pool.images_per_step, thr, per_worker, dt)})
    return out


def reference_arm(args, wl, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on all host threads of this box, same
    metric / config as the fdms arm; rank 0 only (the other ranks exit 0 without work)."""
    if rank != 0:
        return
    blk = reference_block(wl, args.steps, args.warmup, budget_s=150.0)
    v = blk["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic", "config": make_config(wl, world),
        "images_per_step": blk["images_per_step"], "steps_run": blk["steps_run"],
        "cpu_baseline": blk,
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- one workload on the GPU
class Env:
    pass


def measure_workload(name, wl, env, steps, warmup, full, repeats=0, profile=False):
    """Device-resident measurement of one workload on this rank's GPU.  Returns a dict (rank-0 view, max over ranks)."""
    import torch
    import torch.distributed as dist
    from face_detection_multi_scale_b200 import _lib, ops
    from workloads import standin
    dev, world, rank = env.dev, env.world, env.rank
    B, sizes = wl["batch"], wl["sizes"]
    strides, anchors = head_geometry(wl)
    lb_stride = int(max(strides))
    images, heads, shapes = standin.make_batch(wl["h"], wl["w"], sizes, B, wl["faces"], seed=17 + rank, radius=wl["radius"],
                                               strides=strides, anchors_px=anchors)
    geoms = [_lib.letterbox_geometry(wl["h"], wl["w"], s, auto=True, stride=lb_stride) for s in sizes]
    assert [(g.dst_h, g.dst_w) for g in geoms] == [tuple(s) for s in shapes]
    kp = bool(wl.get("kpts"))
    n_sets = 2  # rotate two input sets: nothing of step i is still in L2 for step i+1 (each set >> 126 MB L2 for the batched configs)
    sets = []
    for k in range(n_sets):
        img_d = torch.from_numpy(images if k == 0 else np.roll(images, k, axis=0)).to(dev)
        heads_d = [[torch.from_numpy(l if k == 0 else np.roll(l, k, axis=0)).to(dev) for l in hs] for hs in heads]
        outs = [torch.empty((B, 3, g.dst_h, g.dst_w), dtype=torch.float16, device=dev) for g in geoms]
        plan = ops.PostPlan(heads_d, strides, anchors, [g.dst_h for g in geoms], (wl["h"], wl["w"]), wl["conf"], wl["iou"],
                            cap=wl["cap"], out_cap=wl["cap"], iou_mode=ops.IOU_TV_CUDA, with_kpts=kp, ctx=env.post_ctx[k],
                            small_only=True)  # verified below, like the detector verifies it per batch
        sets.append((img_d, heads_d, outs, plan))

    # The two halves of a step work on independent data (the backbone that links them is out of scope), exactly like
    # consecutive batches of a pipelined deployment (detect_stream, depth 3: pyramid of batch i+2 while batches i and
    # i+1 are post-processed): the pyramid has its own stream and the post-processing chains of consecutive steps
    # alternate between two more (one native context each), so the latency-bound NMS kernels of two batches overlap each
    # other and the HBM-bound pyramid kernel.
    s_pyr, s_posts = env.s_pyr, env.s_post
    pyr_ctx = env.pyr_ctx
    pyr_events = []  # (before, after) CUDA events around sampled pyramid launches of the timed regions, on s_pyr
    post_events = []  # the same around the post-processing chain, on s_post

    def run_pyr(i):
        ops.pyramid(sets[i % n_sets][0], geoms, out_dtype=torch.float16, layout="chw", flip_rb=True, outs=sets[i % n_sets][2], ctx=pyr_ctx)

    def step(i, overlap=True, sample=False):
        plan = sets[i % n_sets][3]
        if overlap:
            with torch.cuda.stream(s_pyr):
                if sample:
                    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ea.record(s_pyr)
                run_pyr(i)
                if sample:
                    eb.record(s_pyr)
                    pyr_events.append((ea, eb))
            s_post = s_posts[i % len(s_posts)]
            with torch.cuda.stream(s_post):
                if sample:
                    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    pa.record(s_post)
                plan.run()
                if sample:
                    pb.record(s_post)
                    post_events.append((pa, pb))
        else:
            run_pyr(i)
            plan.run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def timed(n, overlap, sample=True, use_graph=False):
        """EXACTLY n steps bracketed by barrier + synchronize on both sides, device-timed, max over ranks -> ms."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        barrier()
        e0.record(cur)
        if use_graph:  # graph_steps steps per replay on all streams; the streams join at the end of a replay
            for _ in range(n // env.graph_steps):
                graph.replay()
            if n % env.graph_steps:
                s_pyr.wait_stream(cur)
                for sp in s_posts:
                    sp.wait_stream(cur)
                for i in range(n % env.graph_steps):
                    step(i, overlap, sample=False)
                cur.wait_stream(s_pyr)
                for sp in s_posts:
                    cur.wait_stream(sp)
        else:
            s_pyr.wait_event(e0)
            for sp in s_posts:
                sp.wait_event(e0)
            for i in range(n):
                step(i, overlap, sample=sample and overlap and (i & 3) == 3)  # every 4th launch of an overlapped region is timed
            if overlap:
                cur.wait_stream(s_pyr)
                for sp in s_posts:
                    cur.wait_stream(sp)
        e1.record(cur)
        barrier()
        return allmax(e0.elapsed_time(e1))

    for i in range(max(warmup, 3)):
        step(i, True)
    torch.cuda.synchronize()
    for i in range(3):
        step(i, False)
    torch.cuda.synchronize()

    # capacities / counts of the workload (checked once, outside the timed region).  The plans were built to launch only
    # the single-CTA kernels (what MultiScaleFaceDetector does while its batches stay sparse, verified per batch from
    # the counts it reads back anyway): a workload with a denser segment or image runs the full chain instead.
    plan0 = sets[0][3]
    cnt = plan0.cands.count.cpu().numpy()
    nout = plan0.out.merged_in.cpu().numpy()
    if not ops.PostPlan.small_ok(cnt.max(), nout.max()):
        for st_ in sets:
            st_[3].small_only = False
        for i in range(4):
            step(i, False)
        torch.cuda.synchronize()
        cnt = plan0.cands.count.cpu().numpy()
        nout = plan0.out.merged_in.cpu().numpy()
    nkeep = plan0.cands.n_keep.cpu().numpy()
    if cnt.max() > wl["cap"] or nkeep.min() < 0 or nout.max() > wl["cap"]:
        raise SystemExit("bench.py: workload %s exceeds capacities (count max %d, merged max %d)" % (name, cnt.max(), nout.max()))
    cand_per_step = int(cnt.sum())

    # The steady state (one step per input set, both streams) captured once in a CUDA graph: a replay costs one launch
    # call instead of ~7 through Python, so the timed regions measure the GPU and not the host's launch rate.
    graph = None
    if env.use_graph:
        try:
            cap_stream = torch.cuda.Stream(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap_stream):
                cur = torch.cuda.current_stream(dev)
                s_pyr.wait_stream(cur)
                for sp in s_posts:
                    sp.wait_stream(cur)
                for i in range(env.graph_steps):
                    step(i, True)
                cur.wait_stream(s_pyr)
                for sp in s_posts:
                    cur.wait_stream(sp)
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # capture is an optimisation of the harness: report and fall back to plain launches
            graph = None
            env.graph_error = repr(e)[:200]
    # ---- timed regions: EXACTLY K steps each; the median region is the value ----
    first = timed(steps, True)  # plain launches through Python; its sampled pyramid launches feed `roofline`
    if not repeats:
        repeats = 15 if first < 50.0 else 3
    if profile:
        repeats = 1
    eager = [first] + [timed(steps, True) for _ in range(repeats - 1)]
    regions = [timed(steps, True, use_graph=True) for _ in range(repeats)] if graph is not None else eager
    ms = float(np.median(regions))
    serial = [timed(steps, False, sample=False) for _ in range(1 if profile else min(repeats, 5))]
    ms_serial = float(np.median(serial))
    res = {"value": round(world * B * steps / (ms / 1e3), 1), "ms_per_step": round(ms / steps, 4),
           "repeats": repeats, "region_ms": [round(r, 4) for r in regions],
           "value_min": round(world * B * steps / (max(regions) / 1e3), 1), "value_max": round(world * B * steps / (min(regions) / 1e3), 1),
           "launch_mode": "CUDA graph replay (%d steps per graph, all streams)" % env.graph_steps if graph is not None else "plain launches",
           "eager_ms_per_step": round(float(np.median(eager)) / steps, 4), "graph_error": getattr(env, "graph_error", None),
           "serial_ms_per_step": round(ms_serial / steps, 4), "candidates_per_step": cand_per_step,
           "gpu_launches_per_step": 1 + plan0.launches, "cap": wl["cap"],
           "post_chain": "single-CTA kernels only (every segment and image <= 1024 boxes; checked)" if plan0.small_only else "full chain"}

    # ---- per-stage breakdown + roofline of the dominant kernel (same steps, CUDA events per stage) ----
    stage_names = ["pyramid", "decode_filter", "nms", "merge"]
    K2 = 2 if profile else max(10, min(steps, 50))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stage_names) + 1)] for _ in range(K2)]
    # the host calls of an iteration take about as long as their kernels: park the GPU behind a spin kernel while the
    # whole loop is enqueued, so that the event intervals are kernel time and not launch latency of a starved queue
    ops._ctx(dev.index).set_small_only(plan0.small_only)  # the stages below go through the thread's default context
    if hasattr(torch.cuda, "_sleep"):
        torch.cuda._sleep(int(40e6))
    for i in range(K2):
        img_d, heads_d, outs, plan = sets[i % n_sets]
        c = plan.cands
        evs[i][0].record()
        run_pyr(i)
        evs[i][1].record()
        ops.decode_filter(heads_d, strides, anchors, wl["conf"], wl["cap"], out=c, landmarks=kp)
        evs[i][2].record()
        ops.nms_segments(c.cand, c.count, c.cap, wl["iou"], ops.IOU_TV_CUDA, tie=c.row, keep=c.keep, n_keep=c.n_keep)
        evs[i][3].record()
        ops.merge_scales(c, B, len(sizes), [g.dst_h for g in geoms], (wl["h"], wl["w"]), wl["iou"], wl["cap"], with_kpts=kp, out=plan.out)
        evs[i][4].record()
    torch.cuda.synchronize()
    ops._ctx(dev.index).set_small_only(False)
    stage_ms = {n: float(np.median([evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(K2)])) for j, n in enumerate(stage_names)}

    src_bytes = wl["h"] * wl["w"] * 3
    out_bytes = sum(3 * g.dst_h * g.dst_w * 2 for g in geoms)
    survey_bytes = B * (src_bytes + out_bytes)  # source read once + every scale written once (fp16)
    # source rows any scale's bilinear taps touch (cv2 INTER_LINEAR: fy = (dy + 0.5) * src/dst - 0.5, rows floor(fy) and +1
    # when its weight is not zero): an exact 3:1 scale is a point sample of one row in three, and rows no tap touches are not
    # algorithmic traffic.  All rows for cfg2/cfg3/cfg4 (the 2:1 and 3:2 scales read every row).
    touched = np.zeros(wl["h"], dtype=bool)
    for g in geoms:
        fy = (np.arange(g.unpad_h) + 0.5) * (wl["h"] / g.unpad_h) - 0.5
        y0 = np.floor(fy).astype(np.int64)
        w1 = np.round((fy - y0) * 2048).astype(np.int64)
        touched[np.clip(y0[w1 < 2048], 0, wl["h"] - 1)] = True
        touched[np.clip(y0[w1 > 0] + 1, 0, wl["h"] - 1)] = True
    src_rows = int(touched.sum()) if any(g.resize for g in geoms) else wl["h"]
    pyr_bytes = B * (src_rows * wl["w"] * 3 + out_bytes)
    head_bytes = B * sum(int(np.prod(l.shape[1:])) * 2 for hs in heads for l in hs)
    peak, peak_src = load_peak()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(name, {}).get("pyramid_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    # launch duration of the dominant kernel: CUDA events on ITS stream around sampled launches of the timed (overlapped)
    # regions -- the post-processing kernels share the SMs there; the isolated duration (same kernel, nothing else in
    # flight) is reported next to it
    pyr_ms_timed = float(np.median([a.elapsed_time(b) for a, b in pyr_events])) if pyr_events else stage_ms["pyramid"]
    achieved = pyr_bytes / (pyr_ms_timed * 1e-3) / 1e9
    isolated = pyr_bytes / (stage_ms["pyramid"] * 1e-3) / 1e9
    roofline = {"kernel": "pyramid kernel (fdms_pyramid)", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": pyr_bytes, "source_rows_touched": [src_rows, wl["h"]],
                "survey_8d_bytes_per_launch": survey_bytes, "launch_ms": round(pyr_ms_timed, 4), "launches_timed": len(pyr_events),
                "timing": "CUDA events on the kernel's stream around every 4th launch of the timed regions (post-processing kernels of the "
                          "other stream share the SMs); median",
                "isolated": {"launch_ms": round(stage_ms["pyramid"], 4), "achieved": round(isolated, 1), "frac": round(isolated / peak, 4),
                             "timing": "same kernel with nothing else in flight"},
                "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                "share_of_step": round(stage_ms["pyramid"] / sum(stage_ms.values()), 3),
                "post_chain_in_region_ms": round(float(np.median([a.elapsed_time(b) for a, b in post_events])), 4) if post_events else None}
    decode_gbs = head_bytes / (stage_ms["decode_filter"] * 1e-3) / 1e9
    res.update({"stages_ms": {k: round(v, 4) for k, v in stage_ms.items()},
                "stages_note": ("stages are timed through the separate entry points (decode+filter with landmarks for EVERY candidate); the timed "
                                "step runs fdms_postprocess, which decodes landmarks at emit time for the final detections only") if kp else None,
                "nms_candidates_per_sec": round(world * cand_per_step / (stage_ms["nms"] * 1e-3), 0),
                "roofline": roofline,
                "roofline_decode": {"kernels": "decode_filter_kernel + decode_candidates_kernel", "bound": "hbm", "achieved": round(decode_gbs, 1),
                                    "peak": peak, "unit": "GB/s", "frac": round(decode_gbs / peak, 4),
                                    "launch_ms": round(stage_ms["decode_filter"], 4), "algorithmic_bytes_per_step": head_bytes,
                                    "note": "SURVEY 8(d) algorithmic bytes = the whole raw head (rows x 21 x fp16); the kernels stream only "
                                            "the obj / cls planes and gather the other values per candidate, so this figure is not HBM-bound"}})
    env.keep = (images, heads, shapes, geoms, strides, anchors) if full else None
    del sets
    torch.cuda.empty_cache()
    return res


def h2d_ceiling(env, nbytes=256 << 20, reps=6):
    """Host->device rate of this box with ALL ranks copying at once from pinned memory (GB/s per rank, max-time over ranks)."""
    import torch
    import torch.distributed as dist
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    src.fill_(1)  # first touch
    dst = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    ms = None
    for _ in range(3):  # best of three trials: a ceiling
        if env.world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        ms = t if ms is None else min(ms, t)
    per_rank = nbytes * reps / (ms * 1e-3) / 1e9
    if env.world > 1:
        t = torch.tensor([per_rank], device=env.dev, dtype=torch.float64)
        lst = [torch.zeros_like(t) for _ in range(env.world)]
        dist.all_gather(lst, t)
        rates = [float(x.item()) for x in lst]
    else:
        rates = [per_rank]
    return {"per_gpu_gbs": [round(r, 1) for r in rates], "aggregate_gbs": round(sum(rates), 1),
            "how": "%d x %d MiB pinned host -> device copies per rank, all ranks at once, CUDA events" % (reps, nbytes >> 20)}


def measure_e2e(wl, env, steps):
    """End to end through the public API, host buffers in, host arrays out (rank-local pinned inputs)."""
    import torch
    import torch.distributed as dist
    from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector
    from workloads import standin
    images, heads, shapes, geoms, strides, anchors = env.keep
    dev, world, B, sizes = env.dev, env.world, wl["batch"], wl["sizes"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def timed(fn, n, trials=3):
        """Median of `trials` timed runs of n steps (host-side transients -- page placement, a neighbour's burst on the
        shared PCIe switch -- move a single run by tens of percent)."""
        ts = []
        for _ in range(trials):
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            x0.record()
            fn(n)
            x1.record()
            barrier()
            ts.append(allmax(x0.elapsed_time(x1) / 1e3))
        return float(np.median(ts))

    model = standin.SyntheticHeadModel(strides, anchors)
    for s, hw in enumerate(shapes):
        model.set_heads(hw, heads[s], pin=True)
    det = MultiScaleFaceDetector(model, device=dev, img_sizes=sizes, conf_thres=wl["conf"], iou_thres=wl["iou"], cap=wl["cap"], out_cap=wl["cap"])
    pinned = torch.from_numpy(images).pin_memory()
    res = None
    for _ in range(3):
        res = det.detect_multi_scale_batch(pinned)
    Ke = max(10, min(steps, 100))

    def sync_run(d):
        def f(n):
            for _ in range(n):
                d.detect_multi_scale_batch(pinned)  # ends with the D2H read of the results
        return f
    dt = timed(sync_run(det), Ke)
    # same call with the head maps already resident in HBM (where an on-device backbone leaves them): only the images
    # cross PCIe -- the deployment-shaped number
    model_dev = standin.SyntheticHeadModel(strides, anchors)
    for s_, hw in enumerate(shapes):
        model_dev.set_heads(hw, heads[s_], device=dev)
    det2 = MultiScaleFaceDetector(model_dev, device=dev, img_sizes=sizes, conf_thres=wl["conf"], iou_thres=wl["iou"], cap=wl["cap"], out_cap=wl["cap"])

    # the streaming form of the same API: consecutive batches are pipelined over `depth` streams, so the H2D copy of batch
    # i+1 overlaps the kernels and the D2H read of batch i (every batch still crosses PCIe both ways)
    def stream_run(d, depth=3):
        def f(n):
            for _ in d.detect_stream((pinned for _ in range(n)), depth=depth):
                pass
        return f
    stream_run(det)(4)
    model.h2d_bytes = 0
    dt3 = timed(stream_run(det), Ke)
    h2d_step = int(pinned.numel() + model.h2d_bytes // (3 * Ke))
    stream_run(det2)(4)
    dt2 = timed(stream_run(det2), Ke)
    ceil = h2d_ceiling(env)
    my_rate = ceil["per_gpu_gbs"][env.rank] if env.rank < len(ceil["per_gpu_gbs"]) else ceil["per_gpu_gbs"][0]
    e2e_gbs = h2d_step * Ke / dt3 / 1e9  # per rank
    img_gbs = pinned.numel() * Ke / dt2 / 1e9
    return {"value": round(world * B * Ke / dt3, 1), "unit": "images/s", "steps": Ke, "h2d_bytes_per_step": h2d_step,
            "d2h_bytes_per_step": int(B * wl["cap"] * 7 * 4 + 4 * (B + 5)),  # padded dets + counts, pinned
            "api": "MultiScaleFaceDetector.detect_stream(pinned host uint8 batches, depth=3) -> list of ndarray per batch",
            "h2d_gbs_per_gpu": round(e2e_gbs, 1), "h2d_ceiling": ceil,
            "frac_of_h2d_ceiling": round(world * e2e_gbs / ceil["aggregate_gbs"], 3),
            "sync_call": {"value": round(world * B * Ke / dt, 1), "api": "detect_multi_scale_batch(batch): one synchronous call per batch"},
            "images_only_h2d": {"value": round(world * B * Ke / dt2, 1), "h2d_bytes_per_step": int(pinned.numel()),
                                "h2d_gbs_per_gpu": round(img_gbs, 1), "frac_of_h2d_ceiling": round(world * img_gbs / ceil["aggregate_gbs"], 3),
                                "api": "detect_stream, head maps resident in HBM as an on-device backbone leaves them"},
            "final_detections_per_step": int(sum(len(r) for r in res))}


# ---------------------------------------------------------------------------------------------- main
def main():
    args = parse()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.cap:
        wl["cap"] = args.cap
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return reference_arm(args, wl, rank, world)

    import torch
    import torch.distributed as dist
    from face_detection_multi_scale_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libfdms has no CPU fallback)")
    torch.cuda.set_device(local)
    env = Env()
    env.dev, env.world, env.rank = torch.device("cuda", local), world, rank
    numa_cores = 0
    if world > 1:
        from face_detection_multi_scale_b200 import dist as fdist
        numa_cores = fdist.bind_host_to_gpu(local)  # pinned buffers next to this rank's GPU
        dist.init_process_group("nccl", device_id=env.dev)
    prio = [int(v) for v in os.environ.get("FDMS_BENCH_PRIO", "0,-1").split(",")]  # (pyramid, post): small latency-bound kernels first
    n_post = max(1, int(os.environ.get("FDMS_BENCH_POST_STREAMS", "2")))
    env.s_pyr = torch.cuda.Stream(env.dev, priority=prio[0])
    env.s_post = [torch.cuda.Stream(env.dev, priority=prio[1]) for _ in range(n_post)]
    env.pyr_ctx = _lib.Ctx(local)  # own scratch for the stream that runs the pyramid (a ctx serves one stream)
    env.post_ctx = [_lib.Ctx(local) for _ in range(2)]  # one per input set (= per post-processing stream)
    env.use_graph = not args.no_graph
    env.graph_steps = max(2, min(args.steps, 40)) & ~1

    sampler = ClockSampler(physical_gpu_index(local)) if not args.profile else None
    if sampler:
        sampler.start()
    main_res = measure_workload(args.workload, wl, env, args.steps, args.warmup, full=True, repeats=args.repeats, profile=args.profile)
    clocks = sampler.stop() if sampler else None

    e2e = None
    if not args.no_e2e and not args.profile:
        e2e = measure_e2e(wl, env, args.steps)
    env.keep = None

    configs = {}
    if not args.profile and args.configs != "none":
        names = [c for c in sorted(WORKLOADS) if c != args.workload] if args.configs == "all" else args.configs.split(",")
        for c in names:
            w2 = dict(WORKLOADS[c])
            try:
                r = measure_workload(c, w2, env, max(5, min(args.steps, 20)), 3, full=False, repeats=5)
                configs[c] = {"workload": w2["desc"], "batch_per_gpu": w2["batch"], "value": r["value"], "unit": "images/s",
                              "ms_per_step": r["ms_per_step"], "serial_ms_per_step": r["serial_ms_per_step"], "steps": max(5, min(args.steps, 20)),
                              "repeats": r["repeats"], "stages_ms": r["stages_ms"], "stages_note": r.get("stages_note"), "candidates_per_step": r["candidates_per_step"],
                              "nms_candidates_per_sec": r["nms_candidates_per_sec"], "cap": r["cap"],
                              "roofline": {k: r["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "launch_ms",
                                                                          "algorithmic_bytes_per_launch", "source_rows_touched", "isolated")}}
            except SystemExit as e:
                configs[c] = {"error": str(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:  # N=1 only (measurement contract)
        try:
            cpu = reference_block(wl, steps=args.steps, warmup=2, budget_s=args.cpu_budget)
        except Exception as e:
            cpu = {"error": repr(e)}

    backbone = None
    if rank == 0 and world == 1 and not args.no_backbone and not args.profile:
        # a real random-init reference network on this GPU, in its own process (the reference's `utils` / `models` packages
        # stay out of this one); the main measurements above are finished, the GPU is idle
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "backbone_bench.py")], capture_output=True, text=True, timeout=600)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            backbone = json.loads(lines[-1]) if lines else {"error": (r.stderr or r.stdout)[-400:]}
        except Exception as e:
            backbone = {"error": repr(e)}

    ingest = None
    if rank == 0 and world == 1 and not args.no_backbone and not args.profile:
        # from-files ingest: device JPEG decode against the host decoder, own process
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ingest_bench.py"), "--files", "128"], capture_output=True,
                               text=True, timeout=600)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            ingest = json.loads(lines[-1]) if lines else {"error": (r.stderr or r.stdout)[-400:]}
        except Exception as e:
            ingest = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
            "config": make_config(wl, world),
            "value_is": "median of %d timed regions of exactly %d steps each" % (main_res["repeats"], args.steps),
            "region_ms": main_res["region_ms"], "value_min": main_res["value_min"], "value_max": main_res["value_max"],
            "candidates_per_step": main_res["candidates_per_step"], "candidate_capacity": main_res["cap"],
            "post_chain": main_res["post_chain"],
            "host_cores_bound_to_gpu_numa": numa_cores,
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": args.steps * main_res["repeats"] * main_res["gpu_launches_per_step"],
            "gpu_launches_per_step": main_res["gpu_launches_per_step"],
            "roofline": main_res["roofline"], "cpu_baseline": cpu,
            "streams": "pyramid on one CUDA stream, the post-processing chains of consecutive steps alternating over %d more (independent "
                       "data, overlapped: the pipelining of detect_stream, depth 3)" % len(env.s_post),
            "launch_mode": main_res["launch_mode"], "eager_ms_per_step": main_res["eager_ms_per_step"], "graph_error": main_res["graph_error"],
            "serial_ms_per_step": main_res["serial_ms_per_step"], "stages_ms": main_res["stages_ms"],
            "nms_candidates_per_sec": main_res["nms_candidates_per_sec"],
            "roofline_decode": main_res["roofline_decode"],
            "configs": configs, "backbone": backbone, "ingest": ingest,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
