"""The strip kernel (pyramid_strip.cu: one pass over the source for all exact-factor scales) against the tile kernel
(FDMS_PYR_STRIP=0, pinned against the reference's letterbox by the golden digests) and against the oracle: bit for bit,
over every factor it specialises (1, 3, 5; 2, 4, 6; 3:2), mixed with scales it must leave to the tile kernel, odd batch
sizes, both channel orders, repeated launches (the work counters rearm themselves) and a side stream."""
import os

import numpy as np
import pytest
import torch

from face_detection_multi_scale_b200 import _lib, ops
from oracle import oracle as O
from workloads import synth

pytestmark = pytest.mark.gpu

CASES = [  # (src_h, src_w, sizes, letterbox stride, batch)
    (1080, 1920, (640, 960, 1280), 32, 3),        # cfg2: 3:1 point, 2:1, 3:2 -- all in the strip kernel
    (1080, 1920, (640,), 32, 5),                  # cfg5: one point-sample scale, several periods per strip
    (2160, 3840, (640, 960, 1280, 1536), 64, 2),  # cfg4: 6:1, 4:1, 3:1 in the strip kernel, 2.5:1 in the tile kernel
    (720, 1280, (640, 960, 1280), 32, 2),         # 2:1, 4:3 (tile kernel), 1:1
    (1800, 3200, (640,), 32, 1),                  # 5:1
    (360, 640, (320, 160, 480), 32, 4),           # 2:1, 4:1, 4:3
    (96, 160, (160,), 32, 2),                     # 1:1 with letterbox rows
    (1080, 1920, (1280,), 32, 1),                 # 3:2 alone
]


def _ctx(strip):
    old = os.environ.get("FDMS_PYR_STRIP")
    os.environ["FDMS_PYR_STRIP"] = "2" if strip else "0"  # 2: also launches the strip kernel covers only partly
    try:
        return _lib.Ctx(torch.cuda.current_device())
    finally:
        if old is None:
            del os.environ["FDMS_PYR_STRIP"]
        else:
            os.environ["FDMS_PYR_STRIP"] = old


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("flip", [True, False])
def test_strip_kernel_equals_tile_kernel(case, flip):
    h, w, sizes, stride, B = case
    imgs = np.stack([synth.synth_image(h, w, 40 + b) for b in range(B)])
    src = torch.from_numpy(imgs).cuda()
    geoms = [_lib.letterbox_geometry(h, w, s, auto=True, stride=stride) for s in sizes]
    c_strip, c_tile = _ctx(True), _ctx(False)

    def poisoned():
        return [torch.full((B, 3, g.dst_h, g.dst_w), float("nan"), dtype=torch.float16, device="cuda") for g in geoms]
    ref = ops.pyramid(src, geoms, out_dtype=torch.float16, flip_rb=flip, outs=poisoned(), ctx=c_tile)
    for rep in range(3):  # the strip kernel's work counters must rearm themselves
        got = ops.pyramid(src, geoms, out_dtype=torch.float16, flip_rb=flip, outs=poisoned(), ctx=c_strip)
        torch.cuda.synchronize()
        for s in range(len(sizes)):
            a, b = got[s].view(torch.int16), ref[s].view(torch.int16)
            assert torch.equal(a, b), (case, flip, rep, sizes[s], int((a != b).sum()))
    c_strip.close()
    c_tile.close()


def test_strip_kernel_matches_oracle_and_runs_on_a_side_stream():
    h, w, sizes = 270, 480, (160, 240, 320)  # 3:1, 2:1, 3:2 on a small frame (rows-only padding)
    imgs = np.stack([synth.synth_image(h, w, 7 + b) for b in range(3)])
    geoms = [_lib.letterbox_geometry(h, w, s, auto=True, stride=32) for s in sizes]
    st = torch.cuda.Stream()
    c = _ctx(True)
    with torch.cuda.stream(st):
        src = torch.from_numpy(imgs).cuda()
        outs = ops.pyramid(src, geoms, out_dtype=torch.float16, ctx=c)
    st.synchronize()
    for s, size in enumerate(sizes):
        for b in range(len(imgs)):
            lb, _, _ = O.letterbox(imgs[b], size, auto=True, stride=32)
            want = O.pack_chw(lb, flip=True, dtype=np.float16)
            assert np.array_equal(outs[s][b].cpu().numpy().view(np.uint16), want.view(np.uint16)), (size, b)
    c.close()


def test_strip_kernel_leaves_unaligned_and_api_sources_to_the_tile_kernel():
    """Sources the strip kernel cannot take (row pitch not a multiple of 16 bytes, views with an offset, the api mode's
    virtual square) still give the tile kernel's results through the same entry point."""
    h, w = 1080, 1918  # 5754 bytes per row
    img = synth.synth_image(h, w, 3)
    g = _lib.letterbox_geometry(h, w, 640, auto=True, stride=32)
    (a,) = ops.pyramid(torch.from_numpy(img[None]).cuda(), [g], out_dtype=torch.float16, ctx=_ctx(True))
    lb, _, _ = O.letterbox(img, 640, auto=True, stride=32)
    assert np.array_equal(a[0].cpu().numpy().view(np.uint16), O.pack_chw(lb, True, np.float16).view(np.uint16))
