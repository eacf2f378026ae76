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


def _quit_requested():
    """The reference polls `cv2.waitKey(1) == ord('q')` between frames (utils/datasets.py:226,319); OpenCV builds without
    a GUI raise on that call -- nobody can press a key there."""
    import cv2
    try:
        return cv2.waitKey(1) == ord('q')
    except cv2.error:
        return False


def _chw_rgb(frames, img_size, stride, auto, device, device_out):
    """letterbox(img_size, stride, auto) + BGR->RGB + HWC->CHW of a list of BGR frames: ONE libfdms launch when the
    frames share a shape (the usual case: one camera model, one video), one per frame otherwise.  Returns the list of
    (3, h, w) uint8 images (CUDA tensors, or ndarrays unless `device_out`)."""
    same = all(f.shape == frames[0].shape for f in frames)
    groups = [frames] if same else [[f] for f in frames]
    outs = []
    for grp in groups:
        t = torch.stack([_to_device_u8(f, device) for f in grp]) if len(grp) > 1 else _to_device_u8(grp[0], device)[None]
        g = _lib.letterbox_geometry(t.shape[1], t.shape[2], img_size, auto=auto, stride=stride)
        (o,) = ops.pyramid(t, [g], out_dtype=torch.uint8, layout="chw", flip_rb=True)
        outs.extend(o[i] for i in range(o.shape[0]))
    return outs if device_out else [o.cpu().numpy() for o in outs]


class LoadImages:
    """Drop-in for utils/datasets.py:124-201 (the inference loader of detect.py:61): iterates over the images and the
    frames of the videos of a path / glob / directory and yields `(path, img, img0, cap)` with `img` =
    letterbox(img0, img_size, stride) flipped to RGB and transposed to CHW -- produced by ONE libfdms launch (u8 CHW
    output with the channel flip), bit-identical to the reference's cv2 + numpy sequence.  Every file is decoded once
    on the host (cv2, like the reference).  Same public attributes as the reference's loader (`files`, `video_flag`,
    `nf`, `mode`, `cap`, `count`, `frame`, `nframes`); the walk over the sources is a generator here.
    `device_out=True` yields the CHW image as a CUDA tensor instead of an ndarray; the per-file progress prints of the
    reference are dropped (SURVEY: no prints on the path)."""

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
        self.cap, self.frame, self.nframes, self.count = None, 0, 0, 0
        if clips:  # like the reference (:146-147), the first clip is open from the start (stills of a mixed directory are
            import cv2  # yielded with that capture object)
            self.cap = cv2.VideoCapture(clips[0])
            self.nframes = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))
        assert self.nf > 0, ('No images or videos found in %s. Supported formats are:\nimages: %s\nvideos: %s'
                             % (path, list(IMG_FORMATS), list(VID_FORMATS)))
        self._walk = None

    def _sources(self):
        """(path, BGR frame) of every still, then of every frame of every clip, in file order."""
        import cv2
        for idx, (path, is_clip) in enumerate(zip(self.files, self.video_flag)):
            self.count = idx
            if not is_clip:
                self.mode = 'image'
                img0 = cv2.imread(path)  # BGR
                assert img0 is not None, 'Image Not Found ' + path
                self.count = idx + 1  # the reference has moved on to the next file when it yields a still (:175)
                yield path, img0
                continue
            self.mode = 'video'
            if not (self.cap is not None and idx == self.video_flag.index(True) and self.frame == 0 and self.cap.isOpened()):
                self.cap = cv2.VideoCapture(path)
            self.frame, self.nframes = 0, int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))
            while True:
                ok, img0 = self.cap.read()
                if not ok:
                    break
                self.frame += 1
                yield path, img0
            self.cap.release()
        self.count = self.nf

    def __iter__(self):
        self.count = 0
        self._walk = self._sources()
        return self

    def __next__(self):
        if self._walk is None:
            iter(self)
        path, img0 = next(self._walk)  # StopIteration ends the iteration
        # letterbox + [:, :, ::-1] + transpose(2, 0, 1) + ascontiguousarray (:187-191) in one launch
        (img,) = _chw_rgb([img0], self.img_size, self.stride, False, self.device, self.device_out)  # auto=False (:187)
        return path, img, img0, self.cap

    def __len__(self):
        return self.nf


def clean_str(s):
    """utils/general.py:205-207: characters that do not belong in a window / file name become underscores."""
    import re
    return re.sub(pattern="[|@#!¡·$€%&()=?¿^*;:,¨´><+]", repl="_", string=s)


class LoadStreams:
    """Drop-in for utils/datasets.py:262-340 (`detect.py --source streams.txt | rtsp://... | 0`): one daemon thread per
    source keeps the latest frame (every 4th grabbed frame is retrieved, a failed retrieve blanks the frame, the
    thread sleeps 1 / fps between grabs), `next()` yields `(sources, img (bs, 3, h, w) uint8 RGB, [img0, ...], None)`
    -- the only loader of the reference with batch > 1.  All frames of a tick go through ONE pyramid launch when the
    streams share a shape; `rect` (all letterboxed shapes equal -> minimal padding) is decided like the reference does.
    `device_out=True` keeps `img` on the GPU."""

    def __init__(self, sources='streams.txt', img_size=640, stride=32, device=None, device_out=False):
        import os
        import threading
        import cv2
        self.mode = 'stream'
        self.img_size, self.stride = img_size, stride
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device_out = device_out
        if os.path.isfile(sources) and _suffix(sources) == 'txt':
            with open(sources, 'r') as f:
                sources = [x.strip() for x in f.read().strip().splitlines() if len(x.strip())]
        else:
            sources = [sources]
        self.sources = [clean_str(x) for x in sources]
        self.imgs = [None] * len(sources)
        self.fps = 30.0
        self._caps, self._threads = [], []
        for i, src in enumerate(sources):
            cap = cv2.VideoCapture(int(src) if src.isnumeric() else src)
            assert cap.isOpened(), f'Failed to open {src}'
            self.fps = cap.get(cv2.CAP_PROP_FPS) % 100 or 30.0
            ok, self.imgs[i] = cap.read()  # guarantee first frame
            assert ok, f'Failed to read from {src}'
            th = threading.Thread(target=self.update, args=(i, cap), daemon=True)
            self._caps.append(cap)
            self._threads.append(th)
            th.start()
        shapes = {_lib.letterbox_geometry(x.shape[0], x.shape[1], img_size, auto=True, stride=stride).dst_h * 100000 +
                  _lib.letterbox_geometry(x.shape[0], x.shape[1], img_size, auto=True, stride=stride).dst_w for x in self.imgs}
        self.rect = len(shapes) == 1  # rect inference if all letterboxed shapes are equal (:300-303)

    def update(self, index, cap):
        import time
        n = 0
        while cap.isOpened():
            n += 1
            if not cap.grab():  # end of a file source / a dropped connection
                break
            if n == 4:  # every 4th frame is decoded (:311-315)
                ok, im = cap.retrieve()
                self.imgs[index] = im if ok else self.imgs[index] * 0
                n = 0
            time.sleep(1 / self.fps)

    def __iter__(self):
        self.count = -1
        return self

    def __next__(self):
        self.count += 1
        img0 = self.imgs.copy()
        if _quit_requested():
            raise StopIteration
        outs = _chw_rgb(img0, self.img_size, self.stride, self.rect, self.device, True)
        if not all(o.shape == outs[0].shape for o in outs):
            raise ValueError("LoadStreams: the streams letterbox to different shapes %s and cannot be stacked (the "
                             "reference's np.stack fails the same way)" % sorted({tuple(o.shape) for o in outs}))
        img = torch.stack(outs)
        return self.sources, (img if self.device_out else img.cpu().numpy()), img0, None

    def __len__(self):
        return 0  # endless (:339)


class LoadWebcam:
    """Drop-in for utils/datasets.py:204-259: one local camera (`pipe` numeric: frames are mirrored) or an IP camera
    (every 30th grabbed frame is decoded); yields ('webcam.jpg', img, img0, None)."""

    def __init__(self, pipe='0', img_size=640, stride=32, device=None, device_out=False):
        import cv2
        self.img_size, self.stride = img_size, stride
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device_out = device_out
        self.pipe = int(pipe) if isinstance(pipe, str) and pipe.isnumeric() else pipe
        self.cap = cv2.VideoCapture(self.pipe)
        self.cap.set(cv2.CAP_PROP_BUFFERSIZE, 3)

    def __iter__(self):
        self.count = -1
        return self

    def __next__(self):
        import cv2
        self.count += 1
        if _quit_requested():
            self.cap.release()
            raise StopIteration
        if self.pipe == 0:  # local camera
            ok, img0 = self.cap.read()
            if ok:
                img0 = cv2.flip(img0, 1)
        else:  # IP camera / file: skip to every 30th frame
            ok, img0 = False, None
            for _ in range(30):
                if not self.cap.grab():
                    break
            else:
                ok, img0 = self.cap.retrieve()
        assert ok, f'Camera Error {self.pipe}'
        (img,) = _chw_rgb([img0], self.img_size, self.stride, True, self.device, self.device_out)
        return 'webcam.jpg', img, img0, None

    def __len__(self):
        return 0


def read_images(items, threads=None):
    """Decode a list of image files (paths; already decoded (H, W, 3) uint8 arrays pass through) in parallel on host
    threads with cv2.imread -- the reference's decoder (multi_scale_face_detector.py:84) -- once per image.  The images
    may have different sizes; returns the list of BGR arrays in input order."""
    import concurrent.futures
    import os

    import cv2

    def load(it):
        if isinstance(it, (str, bytes)) or hasattr(it, "__fspath__"):
            img = cv2.imread(os.fsdecode(it))
            if img is None:
                raise ValueError(f"Could not read image: {it}")  # msfd:86
            return img
        img = np.asarray(it)
        if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
            raise ValueError("read_images: images must be (H, W, 3) uint8")
        return img

    items = list(items)
    if not items:
        raise ValueError("read_images: empty list")
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except Exception:
            threads = os.cpu_count() or 1
    if len(items) == 1:
        return [load(items[0])]
    with concurrent.futures.ThreadPoolExecutor(max(1, min(int(threads), len(items)))) as ex:
        return list(ex.map(load, items))


def read_batch_device(paths, device=None, threads=0, out=None):
    """Device-side ingest of a list of same-sized image FILES: the bytes are read on the host, JPEG streams are
    Huffman-decoded on host threads and reconstructed on the GPU (ops.jpeg_decode_batch: IDCT, fancy chroma upsampling,
    YCbCr -> BGR, bit-identical to cv2.imread), everything the device path does not take (PNG, progressive JPEG, EXIF
    rotation, ...) is decoded with cv2.imread and uploaded into its slot.  Returns a (B, H, W, 3) uint8 BGR CUDA tensor whose
    contents are complete when the call returns (the calling thread's current stream is synchronised), and the number of
    images that went through the device path."""
    import os

    import cv2
    paths = [os.fsdecode(p) for p in paths]
    if not paths:
        raise ValueError("read_batch_device: empty batch")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    blobs = []
    for p in paths:
        with open(p, "rb") as f:
            blobs.append(f.read())
    hw = None
    d0 = ops.jpeg_probe(blobs[0])
    if d0 is not None and d0.width > 0 and d0.height > 0:
        hw = (int(d0.height), int(d0.width))
    first = None
    if hw is None:  # not a JPEG: the first image's size comes from the host decoder
        first = cv2.imdecode(np.frombuffer(blobs[0], np.uint8), cv2.IMREAD_COLOR)
        if first is None:
            raise ValueError(f"Could not read image: {paths[0]}")  # msfd:86
        hw = first.shape[:2]
    with torch.cuda.device(dev):
        if out is None or tuple(out.shape) != (len(paths), hw[0], hw[1], 3):
            out = torch.empty((len(paths), hw[0], hw[1], 3), dtype=torch.uint8, device=dev)
        is_jpeg = [b[:2] == b"\xff\xd8" for b in blobs]
        idx = [i for i, j in enumerate(is_jpeg) if j]
        status = np.ones((len(paths),), np.int32)
        if idx:
            if len(idx) == len(paths):
                _, st = ops.jpeg_decode_batch(blobs, hw, out=out, threads=threads)
                status = st
            else:
                sub, st = ops.jpeg_decode_batch([blobs[i] for i in idx], hw, threads=threads, device=dev)
                for k, i in enumerate(idx):
                    status[i] = st[k]
                    if st[k] == 0:
                        out[i].copy_(sub[k])
        for i, s in enumerate(status):
            if s == 0:
                continue
            img = first if (i == 0 and first is not None) else cv2.imread(paths[i])  # exactly the reference's decoder (EXIF etc.)
            if img is None:
                raise ValueError(f"Could not read image: {paths[i]}")
            if img.shape != (hw[0], hw[1], 3):
                raise ValueError("read_batch_device: image %d is %s, the batch is %s" % (i, img.shape, (hw[0], hw[1], 3)))
            out[i].copy_(torch.from_numpy(img))
        torch.cuda.current_stream(dev).synchronize()
    return out, int((status == 0).sum())


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
