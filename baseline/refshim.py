"""Import shim for the upstream reference (TEST INFRASTRUCTURE ONLY).

Makes ``/root/reference`` importable in the build container so that golden
vectors can be generated from the reference's own functions.  The reference
imports matplotlib / seaborn / thop at module import time (utils/general.py:24
-> utils/metrics.py, utils/plots.py:11-15, multi_scale_face_detector.py:7-8);
none of them is installed here and none is on the hot path, so they are stubbed.

Nothing under ``face_detection_multi_scale_b200/`` may import this module.  It is
used by ``tests/golden/make_golden.py`` and ``tests/ref_live_check.py`` (build
container, ``/root/reference``) and by ``baseline/ref_runner.py`` -- the reference
arm of ``bench.py`` and the reference-model GPU tests -- which on the GPU box
finds the byte-for-byte copy ``baseline/_ref/`` (``baseline/install_ref.py``).
"""
import importlib.machinery
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# where the unmodified reference may live: an explicit override, the build container's read-only checkout, or the
# byte-for-byte copy that `baseline/install_ref.py` leaves in the git-ignored baseline/_ref/ (travels to the GPU box)
CANDIDATE_ROOTS = [os.environ.get("FDMS_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
REFERENCE_ROOT = None


def find_root():
    for r in CANDIDATE_ROOTS:
        if r and os.path.isfile(os.path.join(r, "multi_scale_face_detector.py")):
            return r
    return None


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
    global REFERENCE_ROOT
    REFERENCE_ROOT = find_root()
    if REFERENCE_ROOT is None:
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
