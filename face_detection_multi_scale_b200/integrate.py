"""Route (a) of INTEGRATION.md: run UNMODIFIED reference callers (`detect.py`, `test_widerface.py`,
`multi_scale_face_detector.py`, `autoShape`) on libfdms by rebinding the reference's hot-path functions.

    import face_detection_multi_scale_b200.integrate as fi
    fi.monkey_patch()                  # utils.datasets.letterbox, utils.general.non_max_suppression / scale_coords,
                                       # utils.preprocess_yolo_predict.scale_coords_api_approach -> libfdms drop-ins
    model = attempt_load(...)          # the reference's unchanged backbone
    fi.patch_model(model)              # head decode (models/yolo.py:267-308) -> one kernel
    ...
    fi.unpatch()                       # restore the reference's own functions

The reference's callers bind these functions by name at import time (`from utils.general import
non_max_suppression`, multi_scale_face_detector.py:14-15, detect.py:13-15), so rebinding the defining module
alone would not reach them: every already-imported module whose attribute IS the reference function is rebound too.
"""
import sys

from . import datasets as D
from . import general as G
from . import preprocess_yolo_predict as PP
from . import yolo_head

_TARGETS = (("utils.datasets", "letterbox", D.letterbox),
            ("utils.general", "non_max_suppression", G.non_max_suppression),
            ("utils.general", "scale_coords", G.scale_coords),
            ("utils.preprocess_yolo_predict", "scale_coords_api_approach", PP.scale_coords_api_approach))
_saved = []  # (module, attribute, original)


def monkey_patch():
    """Rebind the reference's functions in their defining modules and in every loaded module that imported them.
    The reference package (`utils`, `models`) must be importable.  Returns the number of rebound names."""
    import importlib
    if _saved:
        return len(_saved)
    for mod_name, attr, repl in _TARGETS:
        orig = getattr(importlib.import_module(mod_name), attr)
        for m in list(sys.modules.values()):
            if m is not None and getattr(m, "__dict__", None) is not None and m.__dict__.get(attr) is orig:
                _saved.append((m, attr, orig))
                setattr(m, attr, repl)
    return len(_saved)


def unpatch():
    while _saved:
        m, attr, orig = _saved.pop()
        setattr(m, attr, orig)


def patch_model(model):
    """Replace the decode of the model's detection head (IKeypoint / IDetect / Detect: yolo_head.head_columns)."""
    return yolo_head.patch_head(yolo_head.find_head(model))
