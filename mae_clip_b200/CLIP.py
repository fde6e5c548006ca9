"""Drop-in ``CLIPModel`` / ``cross_entropy`` over the B200 kernels.

Same public surface as ``/root/reference/CLIP.py``: ``CLIPModel(temperature, image_embedding,
text_embedding)`` with attributes ``image_encoder, text_encoder, image_projection,
text_projection, temperature`` and ``forward(batch) -> 0-dim loss``; ``cross_entropy(preds,
targets, reduction)``.  What differs is underneath: lines ``CLIP.py:34-43`` (three BxB matmuls,
softmax, two log-softmaxes on a transposed view, products, sums and their autograd) become one
fused call that never materialises a BxB tensor.
"""
from __future__ import annotations

import torch
from torch import nn

from . import config as CFG
from . import functional as F_b200
from .modules import ImageEncoder, ProjectionHead, TextEncoder


class CLIPModel(nn.Module):
    def __init__(self, temperature=CFG.temperature, image_embedding=CFG.image_embedding,
                 text_embedding=CFG.text_embedding, image_encoder=None, text_encoder=None,
                 gemm_mode=None):
        """``image_encoder`` / ``text_encoder`` (extension): pass ready towers instead of building
        the defaults, e.g. random-init ones offline; ``None`` reproduces ``CLIP.py:17-18``."""
        super().__init__()
        self.image_encoder = image_encoder if image_encoder is not None else ImageEncoder()
        self.text_encoder = text_encoder if text_encoder is not None else TextEncoder()
        self.image_projection = ProjectionHead(embedding_dim=image_embedding, gemm_mode=gemm_mode)
        self.text_projection = ProjectionHead(embedding_dim=text_embedding, gemm_mode=gemm_mode)
        self.temperature = temperature
        self.gemm_mode = gemm_mode

    def forward(self, batch):
        image_features = self.image_encoder(batch["image"])
        text_features = self.text_encoder(
            input_ids=batch["input_ids"], attention_mask=batch["attention_mask"])
        image_embeddings = self.image_projection(image_features)
        text_embeddings = self.text_projection(text_features)
        return F_b200.clip_contrastive_loss(image_embeddings, text_embeddings, self.temperature,
                                            mode=self.gemm_mode or CFG.gemm_mode)


def cross_entropy(preds, targets, reduction="none"):
    """Soft-label CE over the last dim; accepts the non-contiguous ``.T`` views of ``CLIP.py:41``.
    ``'none'`` -> (rows,), ``'mean'`` -> scalar, anything else -> ``None`` (as ``CLIP.py:49-52``)."""
    loss = F_b200.soft_cross_entropy_rows(preds, targets)
    if reduction == "none":
        return loss
    elif reduction == "mean":
        return loss.mean()
