"""B200-native (sm_100a) multi-scale face-detection pre/post-processing -- drop-in for the hot path of
NguyendaiAI-pixta/face_detection_multi_scale.

    from face_detection_multi_scale_b200 import MultiScaleFaceDetector, letterbox, non_max_suppression, scale_coords

Submodules mirror the reference's module names: `datasets` (letterbox), `general` (non_max_suppression,
scale_coords, check_img_size), `preprocess_yolo_predict` (scale_coords_api_approach, preprocess_api_approach),
`yolo_head` (head decode), `multi_scale_face_detector` (MultiScaleFaceDetector), plus `ops` (tensor-level access
to the C ABI), `dist` (image sharding / gather), `evaluation` (bbox_overlaps), `packing` (output tensors).
Importing the package loads nothing native; the first op builds/loads libfdms.so and needs a CUDA device.
"""
__version__ = "0.1.0"

_LAZY = {
    "MultiScaleFaceDetector": ("multi_scale_face_detector", "MultiScaleFaceDetector"),
    "letterbox": ("datasets", "letterbox"),
    "non_max_suppression": ("general", "non_max_suppression"),
    "scale_coords": ("general", "scale_coords"),
    "check_img_size": ("general", "check_img_size"),
    "scale_coords_api_approach": ("preprocess_yolo_predict", "scale_coords_api_approach"),
    "preprocess_api_approach": ("preprocess_yolo_predict", "preprocess_api_approach"),
    "bbox_overlaps": ("evaluation", "bbox_overlaps"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module("." + mod, __name__), attr)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
