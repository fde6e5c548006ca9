"""Oracle for the image side of the data feed ("next" row 4, SURVEY.md section 8 f).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  PARITY UNPINNED against the third-party code:
the reference calls ``A.Normalize(max_pixel_value=255.0)`` (``/root/reference/dataset.py:45-49``) and
``torch.tensor(image).permute(2, 0, 1).float()`` (``dataset.py:34``); albumentations (pinned 1.3.1,
``requirements.txt:4``) is neither vendored under /root/reference nor installed here, so this is a
numpy restatement of its published ``normalize`` function, anchored on the reference's call site.
"""
from __future__ import annotations

import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def normalize_ref(img_hwc_uint8: np.ndarray, mean=MEAN, std=STD, max_pixel_value: float = 255.0) -> np.ndarray:
    """(..., H, W, 3) uint8 -> (..., 3, H, W) float32."""
    mean = np.array(mean, dtype=np.float32) * np.float32(max_pixel_value)
    std = np.array(std, dtype=np.float32) * np.float32(max_pixel_value)
    denominator = np.reciprocal(std, dtype=np.float32)
    img = img_hwc_uint8.astype(np.float32)
    img -= mean
    img *= denominator
    return np.moveaxis(img, -1, -3).copy()     # permute(2, 0, 1)


def token_batch_ref(encoded_captions, indices):
    """Caption half of ``/root/reference/dataset.py:25-28`` followed by the default collate: per index
    ``{key: torch.tensor(values[idx])}``, stacked -> ``{key: (n, L) int64 array}`` (numpy restatement; python-style
    negative indices, IndexError outside the range - exactly what ``values[idx]`` on a list does)."""
    out = {}
    for key in ("input_ids", "attention_mask"):
        values = encoded_captions[key]
        rows = [np.asarray(values[int(i)], dtype=np.int64) for i in indices]
        out[key] = np.stack(rows) if rows else np.zeros((0, len(values[0])), dtype=np.int64)
    return out
