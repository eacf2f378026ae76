"""Output packing of the production driver on libfdms:
`normalize_bbox` (utils/preprocess_yolo_predict.py:26-49), the per-frame lists of
yolov7_face_multi_scale_dataframe_predict.py:699-716 and the -1 padded `[num_frames, max_faces, 4]` tensors of
`create_yolo_json_format` (utils/preprocess_yolo_predict.py:478-571), built on the device from the merged
detections of a batch and read back in one copy."""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops


def pack_frames(dets, n_out, img_w, img_h, max_faces=None):
    """dets (n_img, out_cap, 7) fp32 CUDA + n_out (n_img,) int32 CUDA (ops.Detections) ->
    dict of CUDA tensors: bboxes (n_img, max_faces, 4) f64 normalised, confidence (n_img, max_faces) f64,
    class_indexes / scale_used (n_img, max_faces) int32; padding value -1.  max_faces defaults to max(n_out)
    (one device->host read, like the reference's max over frames, ppyp:493)."""
    if not dets.is_cuda or dets.dtype != torch.float32 or dets.dim() != 3 or dets.shape[2] != 7:
        raise ValueError("pack_frames: dets must be a CUDA fp32 (n_img, out_cap, 7) tensor")
    n_img, out_cap, _ = dets.shape
    if max_faces is None:
        max_faces = int(n_out.clamp(min=0, max=out_cap).max().item()) if n_img else 0
    dev = dets.device
    bboxes = torch.empty((n_img, max_faces, 4), dtype=torch.float64, device=dev)
    conf = torch.empty((n_img, max_faces), dtype=torch.float64, device=dev)
    cls = torch.empty((n_img, max_faces), dtype=torch.int32, device=dev)
    scale = torch.empty((n_img, max_faces), dtype=torch.int32, device=dev)
    c = ops._ctx(dev.index)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rc = _lib.lib().fdms_pack_frames(c.handle, p(dets.contiguous()), p(n_out), n_img, out_cap, max_faces, int(img_w),
                                     int(img_h), p(bboxes), p(conf), p(cls), p(scale),
                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    c.check(rc, "fdms_pack_frames")
    return {"bboxes": bboxes, "confidence": conf, "class_indexes": cls, "scale_used": scale, "max_faces": max_faces}


def create_yolo_json_format(packed, infer_times=None, class_name="face", ckpt_version="yolo_w6_face_v1"):
    """utils/preprocess_yolo_predict.py:478-571 from the packed device tensors (one D2H each).  Returns the same
    dict layout the reference writes; None for an empty batch (:489-490)."""
    bboxes = packed["bboxes"].cpu().numpy()
    n_frames, max_faces = bboxes.shape[:2]
    if n_frames == 0:
        return None
    conf = packed["confidence"].cpu().numpy()
    cls = packed["class_indexes"].cpu().numpy()
    names = np.where(cls >= 0, class_name, "unknown").tolist()
    times = [0.0] * n_frames if infer_times is None else [float(t) for t in infer_times]
    return {"yolo_face_prediction": [
        {"name": "yolo-face-bboxes", "datatype": "FP32", "shape": [n_frames, max_faces, 4], "data": bboxes.tolist()},
        {"name": "yolo-face-confidence", "datatype": "FP32", "shape": [n_frames, max_faces], "data": conf.tolist()},
        {"name": "yolo-face-class_names", "datatype": "BYTES", "shape": [n_frames, max_faces], "data": names},
        {"name": "yolo-face-class_indexes", "datatype": "INT32", "shape": [n_frames, max_faces], "data": cls.tolist()},
        {"name": "yolo-face-class_groups", "datatype": "BYTES", "shape": [n_frames, max_faces], "data": names},
        {"name": "yolo-face-ckpt_version", "datatype": "BYTES", "shape": [n_frames], "data": [ckpt_version] * n_frames},
        {"name": "yolo-face-infer_time", "datatype": "FP32", "shape": [n_frames], "data": times},
        {"name": "yolo-face-total_time", "datatype": "FP32", "shape": [1], "data": [sum(times)]},
    ]}
