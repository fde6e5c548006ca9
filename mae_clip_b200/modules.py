"""Drop-in modules of the hot path.

``ProjectionHead`` keeps the constructor signature, defaults, submodule names and therefore the
``state_dict`` keys of ``/root/reference/modules.py:55-76`` (``projection``, ``fc``,
``layer_norm``), so reference checkpoints (``main.py:118-121`` / ``inference.py:18``) load
unchanged; its forward is one call into the fused CUDA path.  The two encoders are third-party
towers that merely feed the path: they are thin stock-PyTorch
stand-ins with the reference's interface, not part of the accelerated code.
"""
from __future__ import annotations

import torch
from torch import nn

from . import config as CFG
from . import functional as F_b200


class ProjectionHead(nn.Module):
    def __init__(self, embedding_dim, projection_dim=CFG.projection_dim, dropout=CFG.dropout,
                 gemm_mode=None):
        super().__init__()
        # parameter containers only: the arithmetic runs in libmae_clip_b200.so
        self.projection = nn.Linear(embedding_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(projection_dim)
        self.gemm_mode = gemm_mode
        # not a parameter, not a buffer (state_dict keys stay those of the reference): the max-magnitude words of the
        # last forward, from which the next one takes the power-of-two scale of its fp16 operand planes
        self._scale_state = {}

    def forward(self, x, keep_mask=None):
        """``keep_mask`` lets a caller inject the dropout noise (tests: "identical inputs and
        noise"); by default training mode draws it with torch's generator like ``nn.Dropout``."""
        p = self.dropout.p
        if keep_mask is None and self.training and p > 0.0:
            shape = (*x.shape[:-1], self.fc.out_features)
            keep_mask = torch.empty(shape, device=x.device, dtype=torch.uint8).bernoulli_(1.0 - p)
        if not self.training:
            keep_mask = None
        return F_b200.projection_head(
            x, self.projection.weight, self.projection.bias, self.fc.weight, self.fc.bias,
            self.layer_norm.weight, self.layer_norm.bias, keep_mask=keep_mask, p_drop=p,
            eps=self.layer_norm.eps, mode=self.gemm_mode or CFG.gemm_mode, scale_state=self._scale_state)


class ImageEncoder(nn.Module):
    """Image tower stand-in with the reference interface (``modules.py:8-31``): (N,3,H,W) ->
    (N, 2048) pooled features.  Uses ``timm`` when installed, else torchvision's ResNet-50 with the
    classifier removed.  Out of the accelerated scope."""

    def __init__(self, model_name=CFG.model_name, pretrained=CFG.pretrained, trainable=CFG.trainable):
        super().__init__()
        try:
            import timm  # type: ignore
            self.model = timm.create_model(model_name, pretrained, num_classes=0, global_pool="avg")
        except ImportError:
            import torchvision
            if not hasattr(torchvision.models, model_name):
                raise ValueError(f"timm is not installed and torchvision has no '{model_name}'")
            weights = "DEFAULT" if pretrained else None
            net = getattr(torchvision.models, model_name)(weights=weights)
            if hasattr(net, "fc"):
                net.fc = nn.Identity()
            elif hasattr(net, "heads"):
                net.heads = nn.Identity()
            self.model = net
        for p in self.model.parameters():
            p.requires_grad = trainable

    def forward(self, x):
        return self.model(x)


class TextEncoder(nn.Module):
    """DistilBERT tower stand-in with the reference interface (``modules.py:34-51``): CLS hidden
    state, frozen by default.  Out of the accelerated scope."""

    def __init__(self, model_name=CFG.text_encoder_model, pretrained=True, trainable=False):
        super().__init__()
        from transformers import DistilBertConfig, DistilBertModel
        if pretrained:
            self.model = DistilBertModel.from_pretrained(model_name)
        else:
            self.model = DistilBertModel(config=DistilBertConfig())
        for p in self.model.parameters():
            p.requires_grad = trainable
        self.target_token_idx = 0

    def forward(self, input_ids, attention_mask):
        output = self.model(input_ids=input_ids, attention_mask=attention_mask)
        return output.last_hidden_state[:, self.target_token_idx, :]
