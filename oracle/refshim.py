"""Import shim for the upstream reference (TEST INFRASTRUCTURE ONLY).

Makes ``/root/reference`` importable in the build container so that golden
vectors can be generated from the reference's own functions.  The reference
imports matplotlib / seaborn / thop at module import time (utils/general.py:24
-> utils/metrics.py, utils/plots.py:11-15, multi_scale_face_detector.py:7-8);
none of them is installed here and none is on the hot path, so they are stubbed.

Nothing under ``face_detection_multi_scale_b200/`` may import this module, and
nothing that runs on the GPU box may either (``/root/reference`` does not exist
there).  Only ``tests/golden/make_golden.py`` uses it.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FDMS_REFERENCE_ROOT", "/root/reference")


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Dummy(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter(())


def install():
    """Stub plotting deps and put the reference on sys.path. Returns True if usable."""
    if not os.path.isdir(REFERENCE_ROOT):
        return False
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "seaborn", "thop"):
        if name not in sys.modules:
            m = _Dummy(name)
            m.__path__ = []
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return True
