"""Inference-side retrieval ("next" row 3, SURVEY.md section 8 f).

Mirrors ``/root/reference/inference.py:30-47`` (``find_matches``) without the tokenizer download and
the matplotlib display: L2-normalise both embedding sets, ``text @ image.T``, ``torch.topk(n * 5)``,
every fifth hit.  The arithmetic is one streaming pass over the image bank plus an exact radix
select (``mc_similarity_topk``); ties resolve to the lower index.
"""
from __future__ import annotations

import torch

from ._lib import check, cur_stream, lib, ptr, require_cuda, workspace


def similarity_topk(text_embeddings: torch.Tensor, image_embeddings: torch.Tensor, k: int, return_scores: bool = False):
    """(Q, D), (N, D) -> values (Q, k) fp32, indices (Q, k) int64 of the k most similar images per
    query by cosine similarity (``inference.py:42-46``).  ``return_scores`` also returns the (Q, N)
    similarity matrix (``dot_similarity``)."""
    require_cuda(text_embeddings, image_embeddings)
    t = text_embeddings.detach().float().contiguous()
    x = image_embeddings.detach().float().contiguous()
    if t.dim() != 2 or x.dim() != 2 or t.shape[1] != x.shape[1]:
        raise ValueError("similarity_topk expects (Q, D) text and (N, D) image embeddings")
    Q, D = t.shape
    N = x.shape[0]
    if k > N:
        raise RuntimeError(f"selected index k={k} out of range for {N} candidates")  # as torch.topk
    dev = x.device
    vals = torch.empty(Q, k, device=dev, dtype=torch.float32)
    idx = torch.empty(Q, k, device=dev, dtype=torch.int64)
    scores = torch.empty(Q, N, device=dev, dtype=torch.float32) if return_scores else None
    if Q == 0 or N == 0:
        return (vals, idx, scores) if return_scores else (vals, idx)
    with torch.cuda.device(dev):
        nws = lib().mc_similarity_topk_workspace_bytes(Q, N, D)
        ws = workspace(nws, dev)
        check(lib().mc_similarity_topk(ptr(t), Q, ptr(x), N, D, k, ptr(vals), ptr(idx), ptr(scores), ptr(ws), ws.numel(),
                                       cur_stream()), "mc_similarity_topk")
    return (vals, idx, scores) if return_scores else (vals, idx)


def get_image_embeddings(model, loader, device=None):
    """``inference.py:12-28`` minus checkpoint loading: image tower + projection over a loader, eval / no_grad."""
    device = device if device is not None else next(model.parameters()).device
    model.eval()
    out = []
    with torch.no_grad():
        for batch in loader:
            feats = model.image_encoder(batch["image"].to(device))
            out.append(model.image_projection(feats))
    return model, torch.cat(out)


def find_matches(model, image_embeddings, encoded_query, image_filenames, n=9):
    """``inference.py:30-47``: ``encoded_query`` is the tokenizer output for ONE query
    (``{"input_ids": ..., "attention_mask": ...}``, lists or tensors).  Returns the ``n`` matched
    file names (the reference then plots them)."""
    device = image_embeddings.device
    batch = {k: torch.as_tensor(v).to(device) for k, v in encoded_query.items() if k in ("input_ids", "attention_mask")}
    with torch.no_grad():
        text_features = model.text_encoder(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"])
        text_embeddings = model.text_projection(text_features)
    _, indices = similarity_topk(text_embeddings[:1], image_embeddings, n * 5)
    return [image_filenames[i] for i in indices[0, ::5].tolist()]
