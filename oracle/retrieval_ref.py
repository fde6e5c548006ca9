"""Oracle for the inference-side retrieval ("next" row 3, SURVEY.md section 8 f).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Restates ``/root/reference/inference.py:42-46`` op for op (``F.normalize`` on both sides, matmul,
``torch.topk``) on CPU.  Tie order of ``torch.topk`` is unspecified; tests compare values always and
indices where the k-th value is separated from its neighbours.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def similarity_ref(text_embeddings: torch.Tensor, image_embeddings: torch.Tensor, dtype=torch.float32):
    image_n = F.normalize(image_embeddings.to(dtype), p=2, dim=-1)   # inference.py:42
    text_n = F.normalize(text_embeddings.to(dtype), p=2, dim=-1)     # inference.py:43
    return text_n @ image_n.T                                        # inference.py:44


def topk_ref(text_embeddings, image_embeddings, k, dtype=torch.float32):
    return torch.topk(similarity_ref(text_embeddings, image_embeddings, dtype), k)   # inference.py:46


def find_matches_ref(text_embeddings, image_embeddings, image_filenames, n=9):
    _, indices = torch.topk(similarity_ref(text_embeddings, image_embeddings).squeeze(0), n * 5)
    return [image_filenames[idx] for idx in indices[::5]]                            # inference.py:47
