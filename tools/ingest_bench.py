"""From-files ingest: device JPEG decode (host Huffman threads + IDCT / upsampling / colour kernels) against the host
decoder, on a directory of synthetic photo-like 1080p JPEG files written to a temporary directory.

    python tools/ingest_bench.py [--files 128] [--batch 32] [--quality 90]

Prints one JSON line:
  reference_style   cv2.imread once per SCALE and image on one thread (multi_scale_face_detector.py:84 decodes per scale)
  host_all_threads  datasets.read_batch: cv2.imread once per image on all host threads (+ the H2D copy of the pixels)
  device_decode     datasets.read_batch_device: Huffman on all host threads, the rest on the GPU
  detect_files_*    MultiScaleFaceDetector.detect_files end to end (decode + pyramid + post-processing around stand-in heads
                    that are resident on the device), host decode vs device decode; identical detections are asserted
"""
import argparse
import json
import os
import sys
import tempfile
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_multi_scale_b200 import ops  # noqa: E402
from face_detection_multi_scale_b200.datasets import read_batch, read_batch_device  # noqa: E402
from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector  # noqa: E402
from workloads import standin  # noqa: E402


def photo_like(h, w, seed):
    r = np.random.default_rng(seed)
    a = r.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    a = cv2.resize(a, (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip(a.astype(int) + r.integers(-10, 11, (h, w, 3)), 0, 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=128)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--quality", type=int, default=90)
    args = ap.parse_args()
    h, w, sizes, B = 1080, 1920, [640, 960, 1280], args.batch
    dev = torch.device("cuda", 0)
    n_files = max(B, args.files // B * B)
    with tempfile.TemporaryDirectory(prefix="fdms_ingest_") as d:
        paths, nbytes = [], 0
        base = [photo_like(h, w, s) for s in range(8)]
        for i in range(n_files):
            p = os.path.join(d, "img%04d.jpg" % i)
            cv2.imwrite(p, np.roll(base[i % 8], 7 * i, axis=1), [cv2.IMWRITE_JPEG_QUALITY, args.quality])
            nbytes += os.path.getsize(p)
            paths.append(p)
        chunks = [paths[i:i + B] for i in range(0, n_files, B)]
        # reference style: one thread, one imread per scale
        t0 = time.perf_counter()
        for p in paths[:8]:
            for _ in sizes:
                cv2.imread(p)
        ref_style = 8 / (time.perf_counter() - t0)
        # host, all threads, once per image, + upload
        buf = None
        read_batch(chunks[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in chunks:
            buf = read_batch(c, out=buf)
            buf.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        host_all = n_files / (time.perf_counter() - t0)
        # device decode
        out, n_dev = read_batch_device(chunks[0], device=dev)
        assert n_dev == B and np.array_equal(out.cpu().numpy(), read_batch(chunks[0]).numpy())  # bit-identical to cv2
        t0 = time.perf_counter()
        for c in chunks:
            out, _ = read_batch_device(c, device=dev, out=out)
        torch.cuda.synchronize()
        dev_dec = n_files / (time.perf_counter() - t0)
        # detector end to end
        images, heads, shapes = standin.make_batch(h, w, sizes, B, 60, seed=17, radius=1)
        model = standin.SyntheticHeadModel()
        for s, hw in enumerate(shapes):
            model.set_heads(hw, heads[s], device=dev)
        det = MultiScaleFaceDetector(model, device=dev, img_sizes=sizes)
        # the GPU side alone: the same detector over batches already decoded on the device
        list(det.detect_stream([out] * 2))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        list(det.detect_stream([out] * len(chunks)))
        gpu_side = n_files / (time.perf_counter() - t0)
        res = {}
        for mode in (False, True):
            list(det.detect_files(paths[:4 * B], batch_size=B, device_decode=mode))  # warm-up: every buffer slot (depth + 2) exists
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = list(det.detect_files(paths, batch_size=B, device_decode=mode))
            res[mode] = (n_files / (time.perf_counter() - t0), got)
        for (pa, a), (pb, b) in zip(res[False][1], res[True][1]):
            assert pa == pb and a.shape == b.shape and np.array_equal(a, b)
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        threads = os.cpu_count()
    print(json.dumps({
        "files": n_files, "image": "1920x1080 photo-like, 4:2:0, quality %d" % args.quality, "avg_file_kb": round(nbytes / n_files / 1024, 1),
        "host_threads": threads, "batch": B,
        "reference_style_images_per_s": round(ref_style, 1),
        "host_all_threads_images_per_s": round(host_all, 1), "device_decode_images_per_s": round(dev_dec, 1),
        "detect_stream_resident_batches_images_per_s": round(gpu_side, 1),
        "detect_files_host_images_per_s": round(res[False][0], 1), "detect_files_device_decode_images_per_s": round(res[True][0], 1),
        "pixels": "device decode == cv2.imread (asserted); detections identical in both modes (asserted)"}))


if __name__ == "__main__":
    main()
