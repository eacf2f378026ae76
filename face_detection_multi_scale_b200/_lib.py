"""ctypes binding of libfdms.so (include/fdms.h).  No CPU fallback: a missing library is built with
nvcc on first use, and a missing CUDA device makes `ctx()` raise."""
import ctypes as C
import os
import threading

from . import build as _build

FDMS_U8, FDMS_F16, FDMS_F32 = 0, 1, 2
FDMS_CHW, FDMS_HWC = 0, 1
IOU_TV_CPU, IOU_TV_CUDA, IOU_NCNN = 0, 1, 2  # include/fdms.h fdms_iou_mode
MAX_SCALES, MAX_LEVELS, MAX_NMS = 8, 32, 30000
MAX_ANCHORS, MAX_HEAD_NO = 8, 383


class Geom(C.Structure):
    _fields_ = [("dst_h", C.c_int32), ("dst_w", C.c_int32), ("unpad_h", C.c_int32), ("unpad_w", C.c_int32),
                ("top", C.c_int32), ("bottom", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("resize", C.c_int32), ("pad_", C.c_int32), ("ratio_w", C.c_double), ("ratio_h", C.c_double),
                ("dw", C.c_double), ("dh", C.c_double)]


class ScaleOut(C.Structure):
    _fields_ = [("geom", Geom), ("out", C.c_void_p), ("out_img_stride", C.c_int64), ("out_dtype", C.c_int32),
                ("out_layout", C.c_int32)]


class Level(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("ny", C.c_int32), ("nx", C.c_int32),
                ("scale", C.c_int32), ("row_offset", C.c_int32), ("stride", C.c_float), ("anchors", C.c_float * 6)]


class JpegDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32), ("supported", C.c_int32),
                ("h", C.c_int32 * 3), ("v", C.c_int32 * 3), ("blocks_w", C.c_int32 * 3), ("blocks_h", C.c_int32 * 3),
                ("down_w", C.c_int32 * 3), ("down_h", C.c_int32 * 3), ("coef_off", C.c_int64 * 3), ("coef_count", C.c_int64),
                ("qt", (C.c_uint16 * 64) * 3), ("reason", C.c_char * 64)]


class PostDesc(C.Structure):
    _fields_ = [("levels", C.POINTER(Level)), ("n_levels", C.c_int32), ("n_scales", C.c_int32), ("bs", C.c_int32),
                ("cap", C.c_int32), ("conf_thres", C.c_float), ("iou_mode", C.c_int32), ("iou_thres", C.c_double),
                ("in_h_host", C.POINTER(C.c_int32)), ("orig_h", C.c_int32), ("orig_w", C.c_int32),
                ("cand", C.c_void_p), ("cand_row", C.c_void_p), ("count", C.c_void_p), ("keep", C.c_void_p),
                ("n_keep", C.c_void_p), ("dets", C.c_void_p), ("n_out", C.c_void_p), ("out_cap", C.c_int32),
                ("flags", C.c_int32), ("merged_in", C.c_void_p), ("kpts_out", C.c_void_p), ("orig_hw_dev", C.c_void_p)]


_SIGS = {
    "fdms_version": (C.c_int, []),
    "fdms_build_info": (C.c_char_p, []),
    "fdms_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "fdms_destroy": (None, [C.c_void_p]),
    "fdms_last_error": (C.c_char_p, [C.c_void_p]),
    "fdms_scratch_bytes": (C.c_int64, [C.c_void_p]),
    "fdms_letterbox_geometry": (C.c_int, [C.c_int] * 8 + [C.POINTER(Geom)]),
    "fdms_pyramid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.POINTER(ScaleOut), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fdms_decode_dense": (C.c_int, [C.c_void_p, C.POINTER(Level), C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "fdms_decode_dense_head": (C.c_int, [C.c_void_p, C.POINTER(Level), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "fdms_decode_filter": (C.c_int, [C.c_void_p, C.POINTER(Level), C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fdms_decode_filter_ex": (C.c_int, [C.c_void_p, C.POINTER(Level), C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fdms_filter_pred": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fdms_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                           C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdms_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int, C.c_void_p]),
    "fdms_scale_coords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p]),
    "fdms_debug_poison_smem": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fdms_scale_coords_ratio_pad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_double,
                                              C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fdms_scale_coords_api": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]),
    "fdms_merge_scales": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdms_merge_scales_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(C.c_int32), C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdms_weighted_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdms_bbox_overlaps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "fdms_pack_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdms_postprocess": (C.c_int, [C.c_void_p, C.POINTER(PostDesc), C.c_void_p]),
    "fdms_jpeg_probe": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(JpegDesc)]),
    "fdms_jpeg_decode_host": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p]),
    "fdms_jpeg_decode_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_void_p,
                                         C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fdms_set_small_only": (C.c_int, [C.c_void_p, C.c_int]),
    "fdms_set_numerics": (C.c_int, [C.c_void_p, C.c_int]),
    "fdms_get_numerics": (C.c_int, [C.c_void_p]),
    "fdms_scale_img": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "fdms_descale_pred": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()
_tls = threading.local()


def lib():
    """The loaded library (built first if stale)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = _build.build_lib() if os.environ.get("FDMS_NO_BUILD") != "1" else _build.LIB
                if os.environ.get("FDMS_LIB"):  # A/B measurements: another build of the same sources
                    path = os.environ["FDMS_LIB"]
                if not os.path.exists(path):
                    raise RuntimeError("libfdms.so is missing (%s); there is no CPU fallback" % path)
                h = C.CDLL(path)
                for name, (res, args) in _SIGS.items():
                    fn = getattr(h, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = h
    return _lib


class FdmsError(RuntimeError):
    pass


class Ctx:
    """One native context (device scratch).  Not shared between threads: use `ctx(device)`."""

    def __init__(self, device):
        self.device = int(device)
        self._h = C.c_void_p()
        rc = lib().fdms_create(C.byref(self._h), self.device)
        if rc != 0:
            raise FdmsError("fdms_create(device=%d) failed with %d: no usable CUDA device "
                            "(libfdms has no CPU fallback)" % (self.device, rc))

    @property
    def handle(self):
        return self._h

    def check(self, rc, what=""):
        if rc != 0:
            msg = lib().fdms_last_error(self._h).decode("utf-8", "replace")
            exc = ValueError if rc == -1 else FdmsError
            raise exc("%s failed (%d): %s" % (what or "libfdms call", rc, msg))

    def set_numerics(self, mode):
        """fdms_set_numerics: 'exact' (default; the CPU oracle's arithmetic) or 'aten_cuda' (bit-identical to the reference
        running on a CUDA device: libdevice expf in the sigmoid, x * (1 / s) for divisions by a scalar, and -- on fp16 head
        maps, i.e. a half model -- the half roundings of ATen's elementwise kernels in the head decode)."""
        m = {"exact": 0, "aten_cuda": 1, 0: 0, 1: 1}[mode]
        self.check(lib().fdms_set_numerics(self._h, m), "fdms_set_numerics")

    @property
    def numerics(self):
        return "aten_cuda" if lib().fdms_get_numerics(self._h) == 1 else "exact"

    def set_small_only(self, on):
        """fdms_set_small_only: fdms_nms / fdms_merge_scales on this context launch only the single-CTA kernels."""
        self.check(lib().fdms_set_small_only(self._h, int(bool(on))), "fdms_set_small_only")

    def scratch_bytes(self):
        return int(lib().fdms_scratch_bytes(self._h))

    def close(self):
        if self._h:
            lib().fdms_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ctx(device=0, stream=0):
    """Thread-local context for (`device`, `stream`): the handle must not be shared across threads, and its
    scratch buffers and cached tables are ordered by ONE stream, so work enqueued on another CUDA stream of the same
    thread gets a context of its own (`stream` = the cudaStream_t handle as an integer, 0 = the default stream).
    At most 8 per thread are kept; the least recently created one is released beyond that."""
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    key = (int(device), int(stream or 0))
    c = cache.get(key)
    if c is None:
        while len(cache) >= 8:
            cache.pop(next(iter(cache))).close()
        c = cache[key] = Ctx(device)
    return c


def letterbox_geometry(src_h, src_w, new_shape=(640, 640), auto=True, scaleFill=False, scaleup=True, stride=32):
    """utils/datasets.py:873-903 geometry (host arithmetic inside libfdms)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    g = Geom()
    rc = lib().fdms_letterbox_geometry(int(src_h), int(src_w), int(new_shape[0]), int(new_shape[1]), int(bool(auto)),
                                       int(bool(scaleFill)), int(bool(scaleup)), int(stride), C.byref(g))
    if rc != 0:
        raise ValueError("letterbox geometry: bad shape %r -> %r" % ((src_h, src_w), new_shape))
    return g
