"""Multi-GPU plumbing: image sharding (no collective on the hot path) and the optional final gather.

The reference shards inference by item over a spawn Pool, one CUDA context per worker, with no
communication (yolov7_face_multi_scale_dataframe_predict.py:569-596,1115,1147).  Here: one process per
GPU (torchrun), images partitioned contiguously by rank, every rank runs the whole hot path on its shard
with its own libfdms context; the only collective is an OPTIONAL gather of the variable-length results
(two all_gathers: counts, then max-padded rows) over NCCL / NVLink -- gloo on CPU for tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous balanced partition: item i goes to rank i * world // n_items."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(n_items), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard(items, rank=None, world=None):
    """This rank's slice of `items` (list / array).  Defaults to the initialised process group."""
    if rank is None or world is None:
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    a, b = shard_range(len(items), rank, world)
    return items[a:b]


def _comm_device(group=None):
    backend = dist.get_backend(group) if dist.is_initialized() else "gloo"
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def gather_detections(local_dets, n_total, group=None):
    """local_dets: list of (k_i, C) float arrays for this rank's shard (in shard order; np.array([]) allowed).
    Returns the list for ALL n_total images in global order on every rank.  Two collectives per call."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [np.asarray(d, np.float64).reshape(-1, 7) if np.size(d) else np.zeros((0, 7)) for d in local_dets]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = shard_range(n_total, rank, world)
    if len(local_dets) != b - a:
        raise ValueError("rank %d holds %d results for a shard of %d images" % (rank, len(local_dets), b - a))
    dev = _comm_device(group)
    cols = 7
    for d in local_dets:
        if np.size(d):
            cols = np.asarray(d).shape[1]
            break
    rows = [np.asarray(d, np.float64).reshape(-1, cols) if np.size(d) else np.zeros((0, cols)) for d in local_dets]
    per_img = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    counts = torch.zeros((per_img + 1,), dtype=torch.int64)
    counts[:len(rows)] = torch.tensor([r.shape[0] for r in rows], dtype=torch.int64) if rows else 0
    counts[per_img] = cols
    counts = counts.to(dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = torch.stack(all_counts).cpu().numpy()
    cols = int(all_counts[:, per_img].max())
    kmax = int(all_counts[:, :per_img].sum(1).max())
    flat = np.zeros((max(kmax, 1), cols), np.float64)
    if rows and sum(r.shape[0] for r in rows):
        cat = np.concatenate(rows, 0)
        flat[:cat.shape[0], :cat.shape[1]] = cat
    buf = torch.from_numpy(flat).to(dev)
    all_buf = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(all_buf, buf, group=group)
    out = []
    for r in range(world):
        ra, rb = shard_range(n_total, r, world)
        data = all_buf[r].cpu().numpy()
        off = 0
        for i in range(rb - ra):
            k = int(all_counts[r, i])
            out.append(data[off:off + k].copy())
            off += k
    return out


def gather_device(dets, n_out, n_total, group=None):
    """The final gather without leaving the device: `dets` (b, out_cap, C) fp32 and `n_out` (b,) int32 are this rank's
    detections as the merge kernel left them (ops.Detections); returns (all_dets (n_total, kmax, C), all_n (n_total,))
    in global image order on every rank, both device tensors.  Two NCCL collectives (counts, then rows cut to the
    longest list) and one scalar read to size the second; no host copy of the rows, no float64 detour."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        k = int(n_out.max()) if n_out.numel() else 0
        return dets[:, :k].contiguous(), n_out.clone()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = shard_range(n_total, rank, world)
    if dets.shape[0] != b - a or n_out.shape[0] != b - a:
        raise ValueError("rank %d holds %d results for a shard of %d images" % (rank, dets.shape[0], b - a))
    per_img = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    counts = torch.zeros((per_img,), dtype=torch.int32, device=dets.device)
    counts[:b - a] = n_out
    all_counts = torch.empty((world, per_img), dtype=torch.int32, device=dets.device)
    dist.all_gather_into_tensor(all_counts.view(-1), counts, group=group)
    kmax = max(int(all_counts.max()), 1)  # the one host read: sizes the row exchange
    send = torch.zeros((per_img, kmax, dets.shape[2]), dtype=dets.dtype, device=dets.device)
    k_local = min(kmax, dets.shape[1])
    send[:b - a, :k_local] = dets[:, :k_local]
    recv = torch.empty((world,) + tuple(send.shape), dtype=dets.dtype, device=dets.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    keep = torch.cat([torch.arange(r * per_img, r * per_img + (shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0]),
                                   device=dets.device) for r in range(world)])
    return recv.view(world * per_img, kmax, -1)[keep], all_counts.view(-1)[keep]


def detect_sharded(detector, images, gather=True, device_out=False):
    """Run `detector.detect_multi_scale_batch` on this rank's contiguous shard of `images` (B, H, W, 3) and
    optionally gather every image's detections on every rank.  `device_out`: keep everything on the device -- returns
    (dets (n, kmax, 7) fp32, n_out (n,)) tensors (gathered over NCCL when `gather`), no per-image host arrays."""
    n = len(images)
    local = shard(images)
    if device_out:
        dets, n_out = detector.detect_multi_scale_batch(local, device_out=True)
        return gather_device(dets, n_out, n) if gather else (dets, n_out)
    res = detector.detect_multi_scale_batch(local) if len(local) else []
    return gather_detections(res, n) if gather else res


def bind_host_to_gpu(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to the GPU (its NUMA node), so that the pinned
    host buffers allocated afterwards are first-touched next to the GPU's PCIe root.  On an 8-GPU node every rank
    otherwise feeds its GPU across the inter-socket link.  The reference's production driver pins one worker per GPU
    the same way (yolov7_face_multi_scale_dataframe_predict.py:569-596 sets CUDA_VISIBLE_DEVICES per worker).
    Returns the number of cores bound to, or 0 when NVML / affinity are unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis else int(device_index)
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        allowed = set(os.sched_getaffinity(0))
        cores &= allowed
        if not cores or cores == allowed:
            return 0
        os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0
