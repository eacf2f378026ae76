"""Device JPEG ingest on the GPU (csrc/jpeg.cu behind fdms_jpeg_decode_batch): the kernels' pixels are cv2.imdecode's,
bit for bit -- cv2.imread is the reference's own decoder call (multi_scale_face_detector.py:84)."""
import glob
import os

import cv2
import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import ops
from face_detection_multi_scale_b200.datasets import read_batch, read_batch_device
from tests.test_jpeg_ingest import SAMPLING, encoded_cases, photo_like

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_decode(buf):
    return cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)


def test_every_encoded_case_equals_cv2():
    n = 0
    for name, buf in encoded_cases():
        ref = ref_decode(buf)
        out, status = ops.jpeg_decode_batch([buf], ref.shape[:2])
        assert status[0] == 0, (name, status)
        assert np.array_equal(out[0].cpu().numpy(), ref), name
        n += 1
    assert n >= 90


def test_mixed_batch_1080p_one_call():
    """One call, one size, every sampling the device path takes plus streams it must hand back (progressive, 4:1:1)."""
    h, w = 1080, 1920
    bufs, refs = [], []
    for i, (samp, flags) in enumerate([(0x221111, []), (0x211111, []), (0x111111, []), (0x221111, [cv2.IMWRITE_JPEG_RST_INTERVAL, 7]),
                                       (0x221111, [cv2.IMWRITE_JPEG_OPTIMIZE, 1]), (0x221111, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]),
                                       (0x411111, [])]):
        img = photo_like(h, w, 100 + i)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 88, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp] + flags)
        bufs.append(buf.tobytes())
        refs.append(ref_decode(bufs[-1]))
    ok, buf = cv2.imencode(".jpg", photo_like(h, w, 200)[:, :, 1], [cv2.IMWRITE_JPEG_QUALITY, 80])
    bufs.append(buf.tobytes())
    refs.append(ref_decode(bufs[-1]))
    out = torch.full((len(bufs), h, w, 3), 7, dtype=torch.uint8, device="cuda")
    _, status = ops.jpeg_decode_batch(bufs, (h, w), out=out, threads=4)
    assert list(status) == [0, 0, 0, 0, 0, 1, 1, 0], status
    got = out.cpu().numpy()
    for i, s in enumerate(status):
        if s == 0:
            assert np.array_equal(got[i], refs[i]), i
        else:
            assert (got[i] == 7).all()  # untouched: the caller's host decoder fills it
    # a second call on the same context reuses the staging buffers (the previous upload must have completed)
    out2, status2 = ops.jpeg_decode_batch(bufs[:3], (h, w))
    assert np.array_equal(out2.cpu().numpy(), np.stack(refs[:3]))
    # wrong size is an error status, not a crash
    _, st = ops.jpeg_decode_batch(bufs[:1], (h, w - 8))
    assert st[0] < 0


def test_read_batch_device_equals_host_read_batch(tmp_path):
    h, w = 270, 480
    paths = []
    for i in range(6):
        img = photo_like(h, w, 300 + i)
        p = tmp_path / ("f%d.%s" % (i, "png" if i == 2 else "jpg"))
        flags = [cv2.IMWRITE_JPEG_PROGRESSIVE, 1] if i == 4 else []
        assert cv2.imwrite(str(p), img, ([cv2.IMWRITE_JPEG_QUALITY, 90] + flags) if i != 2 else [])
        paths.append(str(p))
    dev, n_dev = read_batch_device(paths)
    assert n_dev == 4  # the PNG and the progressive file went through cv2
    host = read_batch(paths).numpy()
    assert np.array_equal(dev.cpu().numpy(), host)
    photos = sorted(glob.glob(os.path.join(ROOT, "baseline", "_ref", "data", "images", "*.jpg")))
    for f in photos:
        d, n = read_batch_device([f])
        assert n == 1 and np.array_equal(d[0].cpu().numpy(), cv2.imread(f)), f


def test_detect_files_with_device_decode(tmp_path):
    from face_detection_multi_scale_b200.multi_scale_face_detector import MultiScaleFaceDetector
    from workloads import standin
    h, w, sizes, B = 270, 480, [160, 256, 320], 3
    images, heads, shapes = standin.make_batch(h, w, sizes, B, n_faces=20, seed=6, dtype=np.float16)
    model = standin.SyntheticHeadModel()
    for s, hw in enumerate(shapes):
        model.set_heads(hw, heads[s], pin=True)
    det = MultiScaleFaceDetector(model, device="cuda:0", img_sizes=sizes, iou_mode=ops.IOU_TV_CPU)
    paths = []
    for i in range(2 * B):
        p = tmp_path / ("g%d.jpg" % i)
        assert cv2.imwrite(str(p), photo_like(h, w, 400 + i), [cv2.IMWRITE_JPEG_QUALITY, 90])
        paths.append(str(p))
    a = list(det.detect_files(paths, batch_size=B))
    b = list(det.detect_files(paths, batch_size=B, device_decode=True))
    assert [p for p, _ in a] == [p for p, _ in b] == paths
    for (_, x), (_, y) in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)
