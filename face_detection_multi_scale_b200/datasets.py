"""Drop-in for the reference's `utils/datasets.py::letterbox` (utils/datasets.py:873-903) on libfdms.

Same signature, same return value `(img, ratio, (dw, dh))`, numpy HWC uint8 in and out; the resize
(cv2.INTER_LINEAR fixed point, bit exact) and the constant border run in the fused pyramid kernel.
"""
import numpy as np
import torch

from . import _lib, ops


def _to_device_u8(img, device):
    if isinstance(img, torch.Tensor):
        t = img
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError("letterbox: expected an (H, W, 3) uint8 image")
        return t.to(device, non_blocking=True).contiguous()
    a = np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("letterbox: expected an (H, W, 3) uint8 image, got %r %s" % (a.shape, a.dtype))
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


def letterbox(img, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32,
              device=None):
    """utils/datasets.py:873.  Returns (ndarray HWC uint8, (rw, rh), (dw, dh)) like the reference.
    A torch CUDA uint8 image is accepted too and then a CUDA tensor is returned (no host round trip)."""
    color = [int(c) for c in color]
    if len(color) != 3 or min(color) < 0 or max(color) > 255:
        raise ValueError("letterbox: color must be three values in 0..255")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    on_device = isinstance(img, torch.Tensor) and img.is_cuda
    t = _to_device_u8(img, img.device if on_device else dev)
    h, w = t.shape[:2]
    g = _lib.letterbox_geometry(h, w, new_shape, auto=auto, scaleFill=scaleFill, scaleup=scaleup, stride=stride)
    (out,) = ops.pyramid(t[None], [g], out_dtype=torch.uint8, layout="hwc", flip_rb=False, border=int(color[0]))
    out = out[0]
    if color[1] != color[0] or color[2] != color[0]:
        # cv2.copyMakeBorder takes one value per channel (:902).  The kernel fills the border with ONE value (all the
        # detector ever uses is grey); the other two channels of the four border strips are set here, on the device.
        y1, x1 = g.top + g.unpad_h, g.left + g.unpad_w
        for ch in (1, 2):
            plane = out[:, :, ch]
            plane[:g.top].fill_(color[ch])
            plane[y1:].fill_(color[ch])
            plane[g.top:y1, :g.left].fill_(color[ch])
            plane[g.top:y1, x1:].fill_(color[ch])
    ratio, pad = (g.ratio_w, g.ratio_h), (g.dw, g.dh)
    if on_device:
        return out, ratio, pad
    return out.cpu().numpy(), ratio, pad


def letterbox_shape(src_hw, new_shape=640, auto=True, scaleup=True, stride=32):
    """(H, W) of the letterboxed image without touching pixels."""
    g = _lib.letterbox_geometry(src_hw[0], src_hw[1], new_shape, auto=auto, scaleup=scaleup, stride=stride)
    return g.dst_h, g.dst_w


IMG_FORMATS = ('bmp', 'jpg', 'jpeg', 'png', 'tif', 'tiff', 'dng', 'webp', 'mpo')  # utils/datasets.py:29
VID_FORMATS = ('mov', 'avi', 'mp4', 'mpg', 'mpeg', 'm4v', 'wmv', 'mkv')           # utils/datasets.py:30


def _suffix(name):
    return name.rsplit('.', 1)[-1].lower()


def _expand_source(path):
    """Files named by `path`: a glob pattern (recursive), every `*.*` of a directory, or one file -- sorted, absolute,
    exactly the candidates utils/datasets.py:126-134 considers."""
    import glob
    import os
    from pathlib import Path
    spec = str(Path(path).absolute())
    if '*' in spec:
        return sorted(glob.glob(spec, recursive=True))
    if os.path.isdir(spec):
        return sorted(glob.glob(os.path.join(spec, '*.*')))
    if os.path.isfile(spec):
        return [spec]
    raise Exception(f'ERROR: {spec} does not exist')


class LoadImages:
    """Drop-in for utils/datasets.py:124-201 (the inference loader of detect.py:61): iterates over the images and
    videos of a path / glob / directory and yields `(path, img, img0, cap)` with `img` = letterbox(img0, img_size,
    stride, auto=False) flipped to RGB and transposed to CHW -- produced by ONE libfdms launch (u8 CHW output with the
    channel flip), bit-identical to the reference's cv2 + numpy sequence.  The file is decoded once on the host
    (cv2, like the reference).  `device_out=True` yields the CHW image as a CUDA tensor instead of an ndarray; the
    per-file progress prints of the reference are dropped (SURVEY: no prints on the path)."""

    def __init__(self, path, img_size=640, stride=32, device=None, device_out=False):
        self.img_size, self.stride = img_size, stride
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device_out = device_out
        found = _expand_source(path)
        stills = [f for f in found if _suffix(f) in IMG_FORMATS]
        clips = [f for f in found if _suffix(f) in VID_FORMATS]
        self.files = stills + clips                       # images first, then videos (same order as the reference)
        self.video_flag = [False] * len(stills) + [True] * len(clips)
        self.nf = len(self.files)
        self.mode = 'image'
        self.cap = None
        if clips:
            self.new_video(clips[0])
        assert self.nf > 0, ('No images or videos found in %s. Supported formats are:\nimages: %s\nvideos: %s'
                             % (path, list(IMG_FORMATS), list(VID_FORMATS)))

    def __iter__(self):
        self.count = 0
        return self

    def __next__(self):
        import cv2
        if self.count == self.nf:
            raise StopIteration
        path = self.files[self.count]
        if self.video_flag[self.count]:
            self.mode = 'video'
            ret_val, img0 = self.cap.read()
            if not ret_val:
                self.count += 1
                self.cap.release()
                if self.count == self.nf:
                    raise StopIteration
                path = self.files[self.count]
                self.new_video(path)
                ret_val, img0 = self.cap.read()
            self.frame += 1
        else:
            self.count += 1
            img0 = cv2.imread(path)  # BGR
            assert img0 is not None, 'Image Not Found ' + path
        # letterbox(auto=False) + [:, :, ::-1] + transpose(2, 0, 1) + ascontiguousarray (:187-191) in one launch
        t = _to_device_u8(img0, self.device)
        g = _lib.letterbox_geometry(t.shape[0], t.shape[1], self.img_size, auto=False, stride=self.stride)
        (out,) = ops.pyramid(t[None], [g], out_dtype=torch.uint8, layout="chw", flip_rb=True)
        img = out[0] if self.device_out else out[0].cpu().numpy()
        return path, img, img0, self.cap

    def new_video(self, path):
        import cv2
        self.frame = 0
        self.cap = cv2.VideoCapture(path)
        self.nframes = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))

    def __len__(self):
        return self.nf


def read_batch(items, threads=None, out=None):
    """Decode a list of same-sized image files (paths) / already decoded (H, W, 3) uint8 arrays into ONE pinned
    (B, H, W, 3) uint8 BGR host tensor, files in parallel on `threads` host threads (cv2 releases the GIL).  The
    decoder is cv2.imread, i.e. exactly the pixels the reference works on (multi_scale_face_detector.py:84), and each
    file is decoded once for all scales instead of once per scale.  `out` recycles a previous batch buffer."""
    import concurrent.futures
    import os

    import cv2

    def load(it):
        if isinstance(it, (str, bytes)) or hasattr(it, "__fspath__"):
            img = cv2.imread(os.fsdecode(it))
            if img is None:
                raise ValueError(f"Could not read image: {it}")  # msfd:86
            return img
        return np.asarray(it)

    items = list(items)
    if not items:
        raise ValueError("read_batch: empty batch")
    first = load(items[0])
    if first.ndim != 3 or first.shape[2] != 3 or first.dtype != np.uint8:
        raise ValueError("read_batch: images must decode to (H, W, 3) uint8")
    shape = (len(items),) + first.shape
    if out is None or tuple(out.shape) != shape or out.dtype != torch.uint8:
        out = torch.empty(shape, dtype=torch.uint8)
        if torch.cuda.is_available():
            out = out.pin_memory()
    dst = out.numpy()
    dst[0] = first

    def work(i):
        img = load(items[i])
        if img.shape != first.shape:
            raise ValueError("read_batch: image %d is %s, the batch is %s" % (i, img.shape, first.shape))
        dst[i] = img

    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except Exception:
            threads = os.cpu_count() or 1
    if len(items) > 1:
        with concurrent.futures.ThreadPoolExecutor(max(1, min(int(threads), len(items) - 1))) as ex:
            list(ex.map(work, range(1, len(items))))
    return out
