"""Multi-GPU check (torchrun, one process per GPU, NCCL): image-sharded detection + the optional final gather
give every rank the same per-image detections as one GPU processing the whole batch.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_gather_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import dist as fdist, ops  # noqa: E402
from workloads import standin  # noqa: E402
from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
h, w, sizes, B = 270, 480, [160, 256, 320], 7
images, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=20, seed=4, dtype=np.float16)
a, b = fdist.shard_range(B, rank, world)


def detector(lo, hi):
    model = standin.SyntheticHeadModel()
    for s, hw in enumerate(shapes):
        model.set_heads(hw, [l[lo:hi] for l in heads[s]], pin=True)
    return MultiScaleFaceDetector(model, device=dev, img_sizes=sizes, iou_mode=ops.IOU_TV_CPU)


gathered = fdist.detect_sharded(detector(a, b), images, gather=True)
whole = detector(0, B).detect_multi_scale_batch(images)
assert len(gathered) == B
for g, r in zip(gathered, whole):
    r = np.asarray(r).reshape(-1, 7) if np.size(r) else np.zeros((0, 7))
    assert g.shape == r.shape and np.array_equal(g, r)
# the same without leaving the device: fp32 rows and counts gathered over NCCL, only the counts are read by the host
d_all, n_all = fdist.detect_sharded(detector(a, b), images, gather=True, device_out=True)
assert d_all.is_cuda and n_all.is_cuda and d_all.shape[0] == B
for i, r in enumerate(whole):
    r = np.asarray(r).reshape(-1, 7) if np.size(r) else np.zeros((0, 7))
    k = int(n_all[i])
    assert k == len(r) and np.array_equal(d_all[i, :k].cpu().numpy().astype(np.float64), r)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
det_l = detector(a, b)
dl, nl = det_l.detect_multi_scale_batch(images[a:b], device_out=True)
dist.barrier()
ev0.record()
for _ in range(20):
    fdist.gather_device(dl, nl, B)
ev1.record()
torch.cuda.synchronize()
gather_ms = ev0.elapsed_time(ev1) / 20
dist.barrier()
print("rank %d/%d: device gather %.3f ms per call" % (rank, world, gather_ms))
print("rank %d/%d: sharded + NCCL gather == single GPU for %d images (%d detections)" % (rank, world, B, sum(len(g) for g in gathered)))
dist.destroy_process_group()
