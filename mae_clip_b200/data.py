"""The data feed on the GPU ("next" row 4, SURVEY.md section 8 f): images and tokenised captions.

``normalize_images`` replaces the per-sample host work of ``/root/reference/dataset.py:33-34`` and
``:49`` (albumentations ``Normalize(max_pixel_value=255)`` + ``permute(2, 0, 1).float()``): the loader
ships resized uint8 HWC images (a quarter of the bytes of fp32 CHW over PCIe) and one kernel
produces the normalised fp32 NCHW batch.  ``TokenFeed`` replaces the caption half of ``dataset.py:19-31``: the fixed-length
``input_ids`` / ``attention_mask`` tables the tokenizer produced once stay resident in HBM and a batch is a device-side
gather of rows by the sampler's indices.  ``synthetic_batch`` builds the batch dict ``CLIPModel.forward`` consumes
(``CLIP.py:55-58`` shapes) for benchmarks without a dataset.
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


class TokenFeed:
    """Caption side of ``CLIPDataset`` (``/root/reference/dataset.py:19-31``) with the tokenised corpus resident in HBM.

    ``encoded_captions``: what the reference keeps in ``self.encoded_captions`` - the tokenizer's output for the WHOLE
    caption list with ``padding=True`` (every row padded to one length L), i.e. a mapping with ``input_ids`` and
    ``attention_mask`` of N equal-length rows (lists, numpy arrays or tensors).  The tables are uploaded once from pinned
    host memory (asynchronously, on the current stream); ``batch(indices)`` then returns what the reference's
    ``__getitem__`` + default collate produce for those indices: ``{"input_ids": (n, L) int64, "attention_mask": (n, L)
    int64}`` on the device.  ``indices`` may be a list or a host / device int64 tensor; host indices travel through a
    pinned staging buffer.  An index outside [-N, N) raises ``IndexError`` (checked lazily: ``check()`` or the next call).
    """

    def __init__(self, encoded_captions, device="cuda"):
        ids = torch.as_tensor(encoded_captions["input_ids"], dtype=torch.int64)
        mask = torch.as_tensor(encoded_captions["attention_mask"], dtype=torch.int64)
        if ids.dim() != 2 or ids.shape != mask.shape:
            raise ValueError("TokenFeed expects equal-length rows (tokenizer(..., padding=True)) for input_ids / attention_mask")
        self.N, self.L = ids.shape
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("TokenFeed keeps the tokenised corpus on a CUDA device; there is no CPU path")
        self._host = (ids.contiguous().pin_memory(), mask.contiguous().pin_memory())   # kept alive until the copy is done
        self.input_ids = self._host[0].to(self.device, non_blocking=True)
        self.attention_mask = self._host[1].to(self.device, non_blocking=True)
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._stage = None

    def __len__(self):
        return self.N

    def check(self):
        if int(self._bad.item()):
            self._bad.zero_()
            raise IndexError(f"TokenFeed: an index of an earlier batch was outside [-{self.N}, {self.N})")

    def batch(self, indices):
        idx = torch.as_tensor(indices, dtype=torch.int64).reshape(-1)
        n = idx.numel()
        if not idx.is_cuda:
            if self._stage is None or self._stage.numel() < n:
                self._stage = torch.empty(max(n, 1024), dtype=torch.int64).pin_memory()
            self._stage[:n].copy_(idx)
            idx = self._stage[:n].to(self.device, non_blocking=True)
        out_ids = torch.empty(n, self.L, dtype=torch.int64, device=self.device)
        out_mask = torch.empty_like(out_ids)
        if n:
            with torch.cuda.device(self.device):
                check(lib().mc_gather_token_rows(ptr(self.input_ids), ptr(self.attention_mask), self.N, self.L, ptr(idx), n,
                                                 ptr(out_ids), ptr(out_mask), ptr(self._bad), cur_stream()),
                      "mc_gather_token_rows")
        return {"input_ids": out_ids, "attention_mask": out_mask}
