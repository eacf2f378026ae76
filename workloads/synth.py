"""Deterministic synthetic inputs for tests, smoke() and bench.py.

The reference ships no weights and a random-init head yields obj*cls <= ~0.25 everywhere, i.e. zero
candidates at the production thresholds.  Benchmarks and parity tests therefore feed the
post-processing stages *synthetic raw head tensors* -- the (bs, 63, ny, nx) maps an IKeypoint head's 1x1
convolutions would produce (models/yolo.py:271, before the view/permute at :274) -- with a controlled number
of "faces" planted so that the confidence filter, NMS clusters and cross-scale merge all see realistic work.

Everything here is host-side numpy with `numpy.random.default_rng(seed)` (PCG64: bit-reproducible across
machines), so the build container (golden fixtures) and the GPU box generate identical tensors.
"""
import numpy as np

# cfg/*.yaml P5 anchors in pixels (cfg/yolov7s-face.yaml:8-11), strides 8/16/32.
P5_ANCHORS = np.array([[4, 5, 6, 8, 10, 12], [15, 19, 23, 30, 39, 52], [72, 97, 123, 164, 209, 297]],
                      np.float32).reshape(3, 3, 2)
P5_STRIDES = (8, 16, 32)
# cfg/yolov7-w6.yaml:8-12 (P6 head, strides 8/16/32/64)
P6_ANCHORS = np.array([[6, 7, 9, 11, 13, 16], [18, 23, 26, 33, 37, 47], [54, 67, 77, 104, 112, 154],
                       [174, 238, 258, 355, 445, 568]], np.float32).reshape(4, 3, 2)
P6_STRIDES = (8, 16, 32, 64)
NO = 21  # 4 box + obj + 1 cls + 5*(x, y, conf)   (models/yolo.py:222-226 with nc=1, nkpt=5)
NA = 3


def synth_image(h, w, seed):
    """Uniform-random uint8 HWC image."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _logit(p):
    p = np.clip(p, 1e-4, 1 - 1e-4)
    return np.log(p / (1 - p))


def synth_faces(n_faces, img_h, img_w, seed, min_size=12.0, max_size=220.0):
    """Random face boxes (cx, cy, w, h) in ORIGINAL-image pixels, log-uniform sizes."""
    rng = np.random.default_rng(seed)
    size = np.exp(rng.uniform(np.log(min_size), np.log(max_size), n_faces))
    w = size * rng.uniform(0.75, 0.95, n_faces)
    h = size
    cx = rng.uniform(0.05, 0.95, n_faces) * img_w
    cy = rng.uniform(0.05, 0.95, n_faces) * img_h
    return np.stack([cx, cy, w, h], 1)


def synth_head(in_h, in_w, faces_in, seed, strides=P5_STRIDES, anchors_px=P5_ANCHORS, bg_obj=-6.0,
               dtype=np.float32, jitter=0.04, radius=1):
    """Raw head tensors for ONE image at one pyramid scale.

    faces_in: (F,4) boxes (cx,cy,w,h) in the letterboxed input's pixel coordinates.
    Returns a list of per-level arrays shaped (63, ny, nx), channel = anchor*21 + k -- the layout
    IKeypoint reshapes with view(bs, na, no, ny, nx) (models/yolo.py:274).
    """
    rng = np.random.default_rng(seed)
    levels = []
    for li, s in enumerate(strides):
        ny, nx = in_h // s, in_w // s
        t = rng.standard_normal((NA, NO, ny, nx)).astype(np.float32)
        t[:, 4] = bg_obj + rng.standard_normal((NA, ny, nx)).astype(np.float32)
        levels.append(t)
    for f in range(faces_in.shape[0]):
        cx, cy, w, h = faces_in[f]
        # anchors whose size is within a factor ~2 of the face take the detection (best two)
        ratio = np.abs(np.log(np.sqrt(w * h) / np.sqrt(anchors_px[..., 0] * anchors_px[..., 1])))
        order = np.argsort(ratio.reshape(-1))[:2]
        for flat in order:
            li, a = divmod(int(flat), NA)
            s = strides[li]
            t = levels[li]
            ny, nx = t.shape[2], t.shape[3]
            gx0, gy0 = int(cx // s), int(cy // s)
            for gy in range(gy0 - radius, gy0 + radius + 1):
                for gx in range(gx0 - radius, gx0 + radius + 1):
                    if not (0 <= gx < nx and 0 <= gy < ny):
                        continue
                    px, py = (cx / s - gx + 0.5) / 2, (cy / s - gy + 0.5) / 2
                    if not (0.02 < px < 0.98 and 0.02 < py < 0.98):
                        continue
                    aw, ah = anchors_px[li, a]
                    pw, ph = np.sqrt(w / aw) / 2, np.sqrt(h / ah) / 2
                    if not (pw < 0.98 and ph < 0.98):
                        continue
                    j = rng.standard_normal(6) * jitter
                    t[a, 0, gy, gx] = _logit(px) + j[0]
                    t[a, 1, gy, gx] = _logit(py) + j[1]
                    t[a, 2, gy, gx] = _logit(pw) + j[2]
                    t[a, 3, gy, gx] = _logit(ph) + j[3]
                    t[a, 4, gy, gx] = 2.5 + rng.standard_normal() * 1.2
                    t[a, 5, gy, gx] = 3.0 + rng.standard_normal() * 0.8
                    # five landmarks inside the box; x/y are RAW (no sigmoid), conf is a logit
                    kx = cx + w * np.array([-0.2, 0.2, 0.0, -0.15, 0.15]) + rng.standard_normal(5) * jitter * s
                    ky = cy + h * np.array([-0.15, -0.15, 0.05, 0.25, 0.25]) + rng.standard_normal(5) * jitter * s
                    t[a, 6:21:3, gy, gx] = (kx / s - gx + 0.5) / 2
                    t[a, 7:21:3, gy, gx] = (ky / s - gy + 0.5) / 2
                    t[a, 8:21:3, gy, gx] = 2.0 + rng.standard_normal(5)
    return [l.reshape(NA * NO, l.shape[2], l.shape[3]).astype(dtype) for l in levels]


def faces_to_input(faces0, img0_hw, in_hw, api=False):
    """Map original-image face boxes into a scale's letterboxed input coordinates.
    Plain letterbox: scale by r and shift by the pad (utils/datasets.py:880-902); api mode: scale by
    in_h / max(H0, W0), no shift (utils/preprocess_yolo_predict.py:273-290,370)."""
    h0, w0 = img0_hw
    f = faces0.astype(np.float64).copy()
    if api:
        k = in_hw[0] / max(h0, w0)
        f *= k
        return f
    r = min(in_hw[0] / h0, in_hw[1] / w0)
    padx, pady = (in_hw[1] - w0 * r) / 2, (in_hw[0] - h0 * r) / 2
    f *= r
    f[:, 0] += padx
    f[:, 1] += pady
    return f


def synth_pred(rows, n_clusters, per_cluster, seed, extent=(640.0, 640.0), conf_lo=0.3, tie_decimals=None):
    """A decoded `pred` (rows, 21) for ONE image: mostly background rows (obj ~ 0.01) plus clusters of
    near-duplicate boxes -- input for filter+NMS tests that do not need the raw-head layout."""
    rng = np.random.default_rng(seed)
    p = np.zeros((rows, NO), np.float32)
    p[:, 0] = rng.uniform(0, extent[1], rows)
    p[:, 1] = rng.uniform(0, extent[0], rows)
    p[:, 2:4] = rng.uniform(4, 60, (rows, 2))
    p[:, 4] = rng.uniform(0.0, 0.05, rows)
    p[:, 5] = rng.uniform(0.2, 1.0, rows)
    p[:, 6:] = rng.uniform(0, extent[1], (rows, 15))
    n = min(rows, n_clusters * per_cluster)
    idx = rng.permutation(rows)[:n]
    c = rng.integers(0, n_clusters, n)
    ccx = rng.uniform(0.05, 0.95, n_clusters) * extent[1]
    ccy = rng.uniform(0.05, 0.95, n_clusters) * extent[0]
    csz = np.exp(rng.uniform(np.log(10), np.log(160), n_clusters))
    p[idx, 0] = ccx[c] + rng.standard_normal(n) * csz[c] * 0.08
    p[idx, 1] = ccy[c] + rng.standard_normal(n) * csz[c] * 0.08
    p[idx, 2] = csz[c] * np.exp(rng.standard_normal(n) * 0.1)
    p[idx, 3] = csz[c] * 1.2 * np.exp(rng.standard_normal(n) * 0.1)
    p[idx, 4] = rng.uniform(conf_lo ** 0.5, 1.0, n)
    p[idx, 5] = rng.uniform(conf_lo ** 0.5, 1.0, n)
    if tie_decimals is not None:  # force many exact score ties
        p[:, 4] = np.round(p[:, 4], tie_decimals)
        p[:, 5] = np.round(p[:, 5], tie_decimals)
    return p.astype(np.float32)


def letterbox_hw(h, w, size, stride=32):
    """Output (H, W) of utils/datasets.py:letterbox(auto=True) -- integer restatement used to size synthetic heads."""
    r = min(size / h, size / w)
    uw, uh = int(round(w * r)), int(round(h * r))
    dw, dh = (size - uw) % stride, (size - uh) % stride
    dw, dh = dw / 2, dh / 2
    return (uh + int(round(dh - 0.1)) + int(round(dh + 0.1)), uw + int(round(dw - 0.1)) + int(round(dw + 0.1)))


def multi_scale_heads(h, w, img_sizes, n_faces, seed, api=False, strides=P5_STRIDES, anchors_px=P5_ANCHORS,
                      dtype=np.float16, stride_max=32, min_size=10.0, max_size=220.0, radius=1):
    """Raw head tensors of ONE image at every pyramid scale, with the same faces planted at each scale.
    Returns (list over scales of list over levels of (63, ny, nx) arrays, list of model-input (H, W))."""
    faces0 = synth_faces(n_faces, h, w, seed, min_size=min_size, max_size=max_size) if n_faces else np.zeros((0, 4))
    heads, shapes = [], []
    for si, size in enumerate(img_sizes):
        in_hw = (size, size) if api else letterbox_hw(h, w, size, stride_max)
        f_in = faces_to_input(faces0, (h, w), in_hw, api=api)
        heads.append(synth_head(in_hw[0], in_hw[1], f_in, seed * 16 + si, strides=strides, anchors_px=anchors_px,
                                dtype=dtype, radius=radius))
        shapes.append(in_hw)
    return heads, shapes
