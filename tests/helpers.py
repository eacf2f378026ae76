"""Seeded case tables shared by tests/golden/make_golden.py (which runs the reference in the build
container) and the parity tests (which regenerate the same inputs anywhere, GPU box included)."""
import numpy as np

from workloads import synth

LETTERBOX_SMALL = [  # (h, w, new_shape, auto, scaleup, stride, seed)
    (97, 131, 64, True, True, 32, 1), (131, 97, 64, True, True, 32, 2), (37, 53, 96, True, True, 32, 3),
    (37, 53, 96, True, False, 32, 4), (270, 480, 160, True, True, 32, 5), (270, 480, 256, True, True, 32, 6),
    (270, 480, 320, False, True, 32, 7), (64, 64, 64, True, True, 32, 8), (120, 90, 128, False, True, 64, 9),
    (5, 211, 64, True, True, 32, 10), (100, 100, (96, 160), False, True, 32, 11),
]
LETTERBOX_FULL = [  # (h, w, new_shape, auto, stride, seed): SHA-256 digests only
    (1080, 1920, 640, True, 32, 100), (1080, 1920, 960, True, 32, 100), (1080, 1920, 1280, True, 32, 100),
    (2160, 3840, 1536, True, 64, 101), (2160, 3840, 640, True, 64, 101), (720, 1280, 1280, True, 32, 102),
    (480, 640, 1280, True, 32, 103), (1080, 1920, 640, False, 32, 100),
]
API_CASES = [(270, 480, 160, 20), (480, 270, 160, 21), (97, 131, 64, 22), (64, 64, 64, 23), (1080, 1920, 640, 100)]

CROWD_HEADS = {"crowdA": (192, 320, 40, 7), "crowdB": (96, 160, 12, 8)}  # in_h, in_w, n_faces, seed


def crowd_levels(in_h, in_w, n_faces, seed, bs=2):
    """Batch of synthetic raw head tensors: list over levels of (bs, 63, ny, nx) fp32."""
    lv = []
    for b in range(bs):
        faces = synth.synth_faces(n_faces, in_h, in_w, seed + 100 * b, min_size=8, max_size=90)
        lv.append(synth.synth_head(in_h, in_w, faces, seed + 1000 * b))
    return [np.stack([lv[b][i] for b in range(bs)]) for i in range(3)]


E2E_CASES = [  # (name, h, w, img_sizes, n_faces, conf, iou, api, seed)
    ("e2eA", 270, 480, [160, 256, 320], 25, 0.5, 0.5, False, 31),
    ("e2eB", 270, 480, [160, 256, 320], 25, 0.6, 0.3, True, 32),
    ("e2eC", 480, 270, [160, 320], 18, 0.5, 0.3, True, 33),
    ("e2eD", 200, 200, [128], 10, 0.5, 0.5, False, 34),
    ("e2eE", 270, 480, [160, 256, 320], 0, 0.5, 0.5, True, 35),  # nothing found at any scale
]


def letterbox_hw(h, w, size, stride=32):
    """Output (H, W) of utils/datasets.py:letterbox(auto=True) -- pure integer restatement for the tests."""
    r = min(size / h, size / w)
    uw, uh = int(round(w * r)), int(round(h * r))
    dw, dh = (size - uw) % stride, (size - uh) % stride
    dw, dh = dw / 2, dh / 2
    return (uh + int(round(dh - 0.1)) + int(round(dh + 0.1)), uw + int(round(dw - 0.1)) + int(round(dw + 0.1)))


def e2e_heads(h, w, img_sizes, n_faces, api, seed):
    """Per scale: the model-input (H, W) and the synthetic raw head tensors for it."""
    faces0 = synth.synth_faces(n_faces, h, w, seed, min_size=10, max_size=120) if n_faces else np.zeros((0, 4))
    heads, shapes = {}, []
    for si, size in enumerate(img_sizes):
        in_hw = (size, size) if api else letterbox_hw(h, w, size)
        f_in = synth.faces_to_input(faces0, (h, w), in_hw, api=api)
        heads[in_hw] = synth.synth_head(in_hw[0], in_hw[1], f_in, seed * 10 + si)
        shapes.append(in_hw)
    return heads, shapes
