"""Drop-ins for the hot-path functions of the reference's `utils/preprocess_yolo_predict.py`:

    scale_coords_api_approach   utils/preprocess_yolo_predict.py:122-157 (without its debug prints)
    pad_to_square_top_left      utils/preprocess_yolo_predict.py:273-290 (geometry only; the padded square is virtual)
    preprocess_api_approach     utils/preprocess_yolo_predict.py:345-378
"""
import numpy as np
import torch

from . import _lib, ops
from .general import _inplace_on_device


def scale_coords_api_approach(img_input_shape, coords, img0_shape):
    """ppyp:122.  Scales by max(H0, W0) / input_h, clamps to the original image; in place, returns coords."""
    if coords.shape[0] == 0:
        return coords
    return _inplace_on_device(coords, lambda t: ops.scale_coords_api_(img_input_shape, t, img0_shape, False))


def pad_to_square_top_left(h, w):
    """ppyp:273 -- side of the zero-padded square (the image sits at the top-left corner).  libfdms never
    materialises the square: the pyramid kernel reads zeros past the real image."""
    return max(int(h), int(w))


def preprocess_api_approach(image, img_size, stride=32, out_dtype=torch.uint8, device=None):
    """ppyp:345.  `image`: path (decoded with PIL to RGB like the reference) or an (H, W, 3) RGB uint8 array /
    CUDA tensor.  Returns (3, img_size, img_size): a numpy uint8 array like the reference for host inputs, or a
    CUDA tensor (uint8, or fp16/fp32 already divided by 255 with `out_dtype`) for device inputs."""
    if isinstance(image, (str, bytes)) or hasattr(image, "__fspath__"):
        from PIL import Image  # decode is outside the kernel path
        image = np.array(Image.open(image).convert("RGB"))
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    on_device = isinstance(image, torch.Tensor) and image.is_cuda
    t = image if on_device else torch.from_numpy(np.ascontiguousarray(image)).to(dev)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("preprocess_api_approach: expected an (H, W, 3) uint8 RGB image")
    h, w = t.shape[:2]
    s = pad_to_square_top_left(h, w)
    g = _lib.letterbox_geometry(s, s, img_size, auto=False, stride=stride)  # ppyp:370 letterbox(auto=False)
    (out,) = ops.pyramid(t[None].contiguous(), [g], out_dtype=out_dtype, layout="chw", flip_rb=False, virt_hw=(s, s))
    out = out[0]
    return out if on_device else out.cpu().numpy()
