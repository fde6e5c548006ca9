"""Tiny third-party towers for the checkpoint fixture (test infrastructure).  The reference builds its towers through
``timm.create_model`` and ``DistilBertModel(config=DistilBertConfig())`` (``/root/reference/modules.py:17-19,40``); for
a committable checkpoint both calls are pointed at these small stand-ins - the reference's OWN classes (``CLIPModel``,
``ImageEncoder``, ``TextEncoder``, ``ProjectionHead``) are used unmodified around them."""
import torch
from torch import nn

IMG_DIM, TXT_DIM = 48, 32


def tiny_image_model():
    """What the timm stub returns: (N, 3, 32, 32) -> (N, IMG_DIM), `num_classes=0, global_pool="avg"` style."""
    return nn.Sequential(nn.Conv2d(3, 16, 3, stride=2, padding=1), nn.ReLU(), nn.Conv2d(16, IMG_DIM, 3, stride=2, padding=1),
                         nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten())


def tiny_distilbert_config():
    from transformers import DistilBertConfig
    return DistilBertConfig(vocab_size=400, dim=TXT_DIM, n_layers=1, n_heads=2, hidden_dim=64, max_position_embeddings=32,
                            dropout=0.0, attention_dropout=0.0)


class ImageTower(nn.Module):
    """Same attribute layout as the reference's ImageEncoder: the backbone under `.model`."""

    def __init__(self):
        super().__init__()
        self.model = tiny_image_model()

    def forward(self, x):
        return self.model(x)


class TextTower(nn.Module):
    """Same attribute layout as the reference's TextEncoder: DistilBERT under `.model`, CLS token."""

    def __init__(self):
        super().__init__()
        from transformers import DistilBertModel
        self.model = DistilBertModel(config=tiny_distilbert_config())
        self.target_token_idx = 0

    def forward(self, input_ids, attention_mask):
        return self.model(input_ids=input_ids, attention_mask=attention_mask).last_hidden_state[:, self.target_token_idx, :]
