"""Image side of the data feed on the GPU ("next" row 4, SURVEY.md section 8 f).

``normalize_images`` replaces the per-sample host work of ``/root/reference/dataset.py:33-34`` and
``:49`` (albumentations ``Normalize(max_pixel_value=255)`` + ``permute(2, 0, 1).float()``): the loader
ships resized uint8 HWC images (a quarter of the bytes of fp32 CHW over PCIe) and one kernel
produces the normalised fp32 NCHW batch.  ``synthetic_batch`` builds the batch dict
``CLIPModel.forward`` consumes (``CLIP.py:55-58`` shapes) for benchmarks without a dataset.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, cur_stream, lib, ptr, require_cuda

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # albumentations.Normalize defaults
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_images(images_hwc_uint8: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD, max_pixel_value: float = 255.0):
    """(N, H, W, 3) or (H, W, 3) uint8 CUDA tensor -> (N, 3, H, W) fp32, normalised."""
    require_cuda(images_hwc_uint8)
    x = images_hwc_uint8
    if x.dtype != torch.uint8 or x.shape[-1] != 3 or x.dim() not in (3, 4):
        raise ValueError("normalize_images expects uint8 (N, H, W, 3) pixels")
    squeeze = x.dim() == 3
    x = (x.unsqueeze(0) if squeeze else x).contiguous()
    N, H, W, _ = x.shape
    out = torch.empty(N, 3, H, W, device=x.device, dtype=torch.float32)
    if N == 0:
        return out
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    with torch.cuda.device(x.device):
        check(lib().mc_normalize_images(ptr(x), N, H, W, m3, s3, float(max_pixel_value), ptr(out), cur_stream()),
              "mc_normalize_images")
    return out[0] if squeeze else out


def synthetic_batch(batch_size: int, size: int = 224, seq_len: int = 25, device="cuda", generator=None):
    """Random uint8 images + token ids with the shapes of ``CLIP.py:55-58``; images go through
    ``normalize_images`` like real ones would."""
    pix = torch.randint(0, 256, (batch_size, size, size, 3), dtype=torch.uint8, generator=generator)
    ids = torch.randint(5, 300, (batch_size, seq_len), generator=generator)
    return {"image": normalize_images(pix.to(device)), "input_ids": ids.to(device),
            "attention_mask": torch.ones(batch_size, seq_len, dtype=torch.long, device=device)}
