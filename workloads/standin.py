"""Stand-in for the (out-of-scope, unchanged) YOLOv7-face backbone, for tests, smoke() and bench.py.

The reference ships no weights and a random-init network yields no candidates, and the
reference's model code cannot travel to the GPU box.  `SyntheticHeadModel` has the surface the detector
needs from a reference `models.yolo.Model` (`.stride`, `.names`, `.model[-1]` = a head with
nl/na/no/anchor_grid/stride/m, `forward(x, augment=False)`), and its "forward" delivers pre-generated raw
head maps (bs, 63, ny, nx) for the input's (H, W): from pinned host memory with an async H2D copy (the
conservative end-to-end accounting: every byte the post-processing consumes crosses PCIe inside the timed
region) or from device-resident tensors.
"""
import numpy as np
import torch

from . import synth


class _Head(torch.nn.Module):
    """Attribute-compatible with models/yolo.py::IKeypoint after its 1x1 convolutions (identity convs)."""

    def __init__(self, strides, anchors_px):
        super().__init__()
        self.nl = len(strides)
        self.na = 3
        self.no = synth.NO
        self.nc = 1
        self.nkpt = 5  # no m_kpt branch: the synthetic maps already hold all 63 channels (yolo_head.head_convs)
        self.stride = torch.tensor([float(s) for s in strides])
        a = torch.as_tensor(np.asarray(anchors_px, np.float32)).view(self.nl, -1, 2)
        self.register_buffer("anchor_grid", a.clone().view(self.nl, 1, -1, 1, 1, 2))
        self.m = torch.nn.ModuleList(torch.nn.Identity() for _ in strides)

    def forward(self, x):  # replaced by yolo_head.patch_head
        raise RuntimeError("stand-in head must be patched by yolo_head.patch_head")


class SyntheticHeadModel(torch.nn.Module):
    def __init__(self, strides=synth.P5_STRIDES, anchors_px=synth.P5_ANCHORS):
        super().__init__()
        self.model = torch.nn.Sequential(_Head(strides, anchors_px))
        self.stride = torch.tensor([float(s) for s in strides])
        self.names = ["face"]
        self._heads = {}  # (H, W) -> list of per-level tensors (pinned host or device)
        self.h2d_bytes = 0

    def half(self):  # the maps carry their own dtype; nothing to convert
        return self

    def set_heads(self, in_hw, levels, pin=False, device=None):
        """levels: list of (bs, 63, ny, nx) numpy arrays / tensors for model inputs of size in_hw."""
        ts = []
        for l in levels:
            t = torch.from_numpy(np.ascontiguousarray(l)) if isinstance(l, np.ndarray) else l
            if device is not None:
                t = t.to(device)
            elif pin:
                t = t.pin_memory()
            ts.append(t)
        self._heads[(int(in_hw[0]), int(in_hw[1]))] = ts

    def forward(self, x, augment=False):
        levels = self._heads[(int(x.shape[2]), int(x.shape[3]))]
        out = []
        for t in levels:
            if t.shape[0] != x.shape[0]:
                raise ValueError("stand-in heads were generated for batch %d, got %d" % (t.shape[0], x.shape[0]))
            if not t.is_cuda:
                self.h2d_bytes += t.numel() * t.element_size()
                t = t.to(x.device, non_blocking=True)
            out.append(t)
        return self.model[-1](out)


def make_batch(h, w, img_sizes, batch, n_faces, seed, api=False, dtype=np.float16, strides=synth.P5_STRIDES,
               anchors_px=synth.P5_ANCHORS, radius=1, min_size=10.0, max_size=220.0):
    """Synthetic workload: images (B, H, W, 3) u8 and, per scale, per level, stacked raw heads (B, 63, ny, nx).
    Returns (images, heads[scale][level], shapes[scale])."""
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, (batch, h, w, 3), dtype=np.uint8)
    per_img = [synth.multi_scale_heads(h, w, img_sizes, n_faces, seed * 1000 + b, api=api, dtype=dtype, strides=strides,
                                       anchors_px=anchors_px, stride_max=int(max(strides)), radius=radius,
                                       min_size=min_size, max_size=max_size)
               for b in range(batch)]
    shapes = per_img[0][1]
    heads = [[np.stack([per_img[b][0][s][l] for b in range(batch)]) for l in range(len(strides))]
             for s in range(len(img_sizes))]
    return images, heads, shapes
