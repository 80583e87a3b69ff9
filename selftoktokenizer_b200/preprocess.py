"""Host-side image boundary of the reference's driver script (test.py:27-31, 45-47): resize + centre crop + NormalizeToTensor on the
way in, `save_image` on the way out.  PIL only (torchvision is optional in a serving image); the arithmetic follows
torchvision.transforms.Resize(int) / CenterCrop(int) on PIL inputs and torchvision.utils.save_image for one image, and
tests/test_host_and_abi.py checks it against torchvision when that is importable.  SURVEY 8a13 / 8f4: stays on the host."""
from __future__ import annotations

from typing import Iterable, Union

import numpy as np
import torch

from .pipeline import NormalizeToTensor


def resize_center_crop(img, size: int):
    """PIL image -> PIL image [size, size]: the smaller edge is resized to `size` (bilinear, PIL's antialiasing reducer, the
    longer edge int(size * long / short) -- transforms.Resize(size)), then the centre [size, size] window with torchvision's
    rounding (transforms.CenterCrop(size): top = int(round((h - size) / 2.0)))."""
    from PIL import Image
    w, h = img.size
    if (w <= h and w != size) or (h <= w and h != size):
        if w <= h:
            nw, nh = size, int(size * h / w)
        else:
            nw, nh = int(size * w / h), size
        img = img.resize((nw, nh), Image.BILINEAR)
    w, h = img.size
    top, left = int(round((h - size) / 2.0)), int(round((w - size) / 2.0))
    return img.crop((left, top, left + size, top + size))


def load_images(paths_or_images: Iterable[Union[str, "object"]], size: int) -> torch.Tensor:
    """test.py:33-35: [transform(Image.open(p)) for p in paths] stacked -> float32 [B, 3, size, size] in [-1, 1] (host)."""
    from PIL import Image
    tf = NormalizeToTensor()
    out = []
    for p in paths_or_images:
        img = Image.open(p) if isinstance(p, (str, bytes)) or hasattr(p, "__fspath__") else p
        out.append(tf(resize_center_crop(img.convert("RGB"), size)))
    return torch.stack(out)


def to_uint8_hwc(image: torch.Tensor) -> np.ndarray:
    """One [3, H, W] image in [0, 1] (what decoding() returns after norm_ip) -> uint8 [H, W, 3], torchvision.utils.save_image's
    quantisation: mul(255).add_(0.5).clamp_(0, 255).to(uint8)."""
    x = image.detach().to(torch.float32).cpu()
    return x.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()


def save_image(image: torch.Tensor, path: str) -> None:
    from PIL import Image
    Image.fromarray(to_uint8_hwc(image)).save(path)
