"""SURVEY 8f "next" rows built so far: the WIDER-FACE evaluation IoU matrix (rank 4) and the padded per-frame
output tensors of the production driver (rank 2).  Oracle vs reference golden on CPU; CUDA vs golden on GPU."""
import numpy as np
import pytest

from oracle import oracle as O


def test_oracle_bbox_overlaps_golden(golden):
    g = golden("bbox")
    for i in range(5):
        assert np.array_equal(O.bbox_overlaps(g["b%d" % i], g["q%d" % i]), g["o%d" % i])


def _pack_ref(dets_list, w, h):
    """numpy restatement of normalize_bbox (ppyp:26-49) + the -1 padding of create_yolo_json_format (ppyp:493-520)."""
    m = max([len(d) for d in dets_list] + [0])
    bb = -np.ones((len(dets_list), m, 4)); cf = -np.ones((len(dets_list), m)); ci = -np.ones((len(dets_list), m), np.int32)
    for f, d in enumerate(dets_list):
        for k, r in enumerate(d):
            bb[f, k] = [float(r[0]) / w, float(r[1]) / h, float(r[2]) / w, float(r[3]) / h]
            cf[f, k] = float(r[4]); ci[f, k] = int(r[5])
    return bb, cf, ci


def test_pack_restatement_matches_reference(golden):
    g = golden("pack")
    w, h = g["wh"]
    dets = [g["dets%d" % i] for i in range(len(g["counts"]))]
    bb, cf, ci = _pack_ref(dets, int(w), int(h))
    assert list(g["shape"]) == list(bb.shape)
    assert np.array_equal(bb, g["bboxes"]) and np.array_equal(cf, g["confidence"]) and np.array_equal(ci, g["class_indexes"])


@pytest.mark.gpu
def test_gpu_bbox_overlaps(golden):
    import torch
    from face_detection_multi_scale_b200.evaluation import bbox_overlaps
    g = golden("bbox")
    for i in range(5):
        out = bbox_overlaps(g["b%d" % i], g["q%d" % i])
        assert out.dtype == np.float64 and out.shape == g["o%d" % i].shape
        assert np.array_equal(out, g["o%d" % i])
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 2000, (3000, 4)); a[:, 2:] += a[:, :2]
    q = rng.uniform(0, 2000, (700, 4)); q[:, 2:] += q[:, :2]
    d = bbox_overlaps(torch.from_numpy(a).cuda(), torch.from_numpy(q).cuda())
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), O.bbox_overlaps(a, q))


@pytest.mark.gpu
def test_gpu_pack_frames_and_json(golden):
    import torch
    from face_detection_multi_scale_b200 import packing
    g = golden("pack")
    w, h = int(g["wh"][0]), int(g["wh"][1])
    counts = g["counts"]
    cap = 16
    dets = np.zeros((len(counts), cap, 7), np.float32)
    for i, n in enumerate(counts):
        dets[i, :n] = g["dets%d" % i]
    packed = packing.pack_frames(torch.from_numpy(dets).cuda(), torch.from_numpy(counts.astype(np.int32)).cuda(), w, h)
    assert packed["max_faces"] == counts.max()
    assert np.array_equal(packed["bboxes"].cpu().numpy(), g["bboxes"])
    assert np.array_equal(packed["confidence"].cpu().numpy(), g["confidence"])
    assert np.array_equal(packed["class_indexes"].cpu().numpy(), g["class_indexes"])
    js = packing.create_yolo_json_format(packed, infer_times=[0.01, 0.02, 0.03, 0.04])["yolo_face_prediction"]
    assert js[0]["shape"] == list(g["shape"]) and js[2]["data"][1] == ["unknown"] * counts.max()
    assert [e["name"] for e in js] == ["yolo-face-bboxes", "yolo-face-confidence", "yolo-face-class_names",
                                       "yolo-face-class_indexes", "yolo-face-class_groups", "yolo-face-ckpt_version",
                                       "yolo-face-infer_time", "yolo-face-total_time"]
