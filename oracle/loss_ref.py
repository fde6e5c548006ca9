"""Oracle for the contrastive soft-target loss.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Two independent statements of the same maths:

``clip_loss_ref``       op-for-op restatement of the reference forward
                        (``/root/reference/CLIP.py:34-43`` and the helper at
                        ``CLIP.py:46-52``) on plain torch CPU tensors; autograd
                        supplies the backward exactly as ``main.py:58`` does.
``clip_loss_closed_form`` numpy float64 closed form of loss and gradients
                        (SURVEY.md section 8 row L6), with no autograd; it is the
                        tie-breaker when fp32 rounding of the two paths differs.
"""
from __future__ import annotations

import numpy as np
import torch


def soft_cross_entropy_ref(preds: torch.Tensor, targets: torch.Tensor, reduction: str = "none"):
    """Soft-label CE over the last dim of a 2-D tensor.

    Follows ``/root/reference/CLIP.py:46-52``: per-row ``-(targets * log_softmax(preds)).sum(1)``;
    ``'none'`` -> vector, ``'mean'`` -> scalar, any other string -> ``None`` (the
    reference falls off the end of the function).
    """
    logp = torch.log_softmax(preds, dim=-1)
    per_row = (-targets * logp).sum(1)
    if reduction == "none":
        return per_row
    if reduction == "mean":
        return per_row.mean()
    return None


def clip_loss_ref(image_emb: torch.Tensor, text_emb: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
    """Scalar loss from (B, D) embeddings, reference op order (``CLIP.py:34-43``).

    logits are text @ image^T *divided* by temperature (``:34``); the target
    similarities are *multiplied* by it (``:38``); the targets are NOT detached
    (``:37-39``), so autograd differentiates through them; the second CE runs
    on the transposed views (``:41``).
    """
    logits = (text_emb @ image_emb.T) / temperature
    sim_i = image_emb @ image_emb.T
    sim_t = text_emb @ text_emb.T
    targets = torch.softmax((sim_i + sim_t) / 2 * temperature, dim=-1)
    loss_t = soft_cross_entropy_ref(logits, targets, "none")
    loss_i = soft_cross_entropy_ref(logits.T, targets.T, "none")
    return ((loss_i + loss_t) / 2.0).mean()


def clip_loss_fwd_bwd_ref(image_emb, text_emb, temperature=1.0, dtype=torch.float32):
    """(loss, dI, dT) through torch autograd on CPU in ``dtype``."""
    I = torch.as_tensor(image_emb).detach().to(dtype).clone().requires_grad_(True)
    T = torch.as_tensor(text_emb).detach().to(dtype).clone().requires_grad_(True)
    loss = clip_loss_ref(I, T, temperature)
    loss.backward()
    return loss.detach(), I.grad.detach(), T.grad.detach()


def _lse(a: np.ndarray, axis: int) -> np.ndarray:
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def clip_loss_closed_form(image_emb, text_emb, temperature: float = 1.0, grad_loss: float = 1.0):
    """float64 closed form of loss, dI, dT and the row statistics the kernels keep.

    With S = T I^T / tau, Z = (I I^T + T T^T) tau / 2, P = softmax_row(Z),
    r = rowLSE(S), c = colLSE(S):
      loss = -(1/2B) sum_ij P_ij (2 S_ij - r_i - c_j)
      dS   = [softmax_row(S) + softmax_col(S) * colsum(P) - 2 P] / (2B)
      G    = -(2 S - r_i - c_j) / (2B);  dZ = P * (G - rowsum(G * P))
      dI   = dS^T T / tau + (tau/2) (dZ + dZ^T) I
      dT   = dS I / tau   + (tau/2) (dZ + dZ^T) T
    (derived from ``CLIP.py:34-43``.)
    """
    I = np.asarray(image_emb, dtype=np.float64)
    T = np.asarray(text_emb, dtype=np.float64)
    B = I.shape[0]
    tau = float(temperature)
    S = (T @ I.T) / tau
    Z = (I @ I.T + T @ T.T) * (tau / 2.0)
    r = _lse(S, 1)
    c = _lse(S, 0)
    rz = _lse(Z, 1)
    P = np.exp(Z - rz[:, None])
    G = -(2.0 * S - r[:, None] - c[None, :]) / (2.0 * B)
    g = (G * P).sum(1)
    loss = g.sum()
    q = P.sum(0)
    dS = (np.exp(S - r[:, None]) + np.exp(S - c[None, :]) * q[None, :] - 2.0 * P) / (2.0 * B)
    dZ = P * (G - g[:, None])
    dZs = dZ + dZ.T
    dI = (dS.T @ T) / tau + (tau / 2.0) * (dZs @ I)
    dT = (dS @ I) / tau + (tau / 2.0) * (dZs @ T)
    stats = {"row_lse_s": r, "col_lse_s": c, "row_lse_z": rz, "row_g": g, "col_sum_p": q}
    return loss, dI * grad_loss, dT * grad_loss, stats


def make_embeddings(batch: int, dim: int = 256, seed: int = 0, scale: float = 1.0, dtype=torch.float32):
    """Synthetic embeddings with the distribution ProjectionHead emits: LayerNorm'd
    gaussian rows (mean 0, var 1, norm sqrt(dim)); ``scale`` < 1 gives the
    soft-target regime where P is far from one-hot (SURVEY.md section 7 hard part b)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, dim, generator=g, dtype=torch.float32)
    x = torch.nn.functional.layer_norm(x, (dim,)) * scale
    return x.to(dtype)


def tile_relevance(image_emb, text_emb, temperature: float = 1.0, tile: int = 128, theta_log2: float = 44.0):
    """Exact tile relevance of the soft targets (checker for the CUDA engine's tile flags): entry (I, J) is True when
    some P_ij >= 2^-theta or some P_ji >= 2^-theta with i in row block I and j in column tile J, where
    P = softmax_row(Z), Z = (I I^T + T T^T) tau / 2 (``CLIP.py:35-39``).  The engine may flag MORE tiles (its test is a
    bound), never fewer."""
    import numpy as np
    I = np.asarray(image_emb, dtype=np.float64)
    T = np.asarray(text_emb, dtype=np.float64)
    Z = (I @ I.T + T @ T.T) * (temperature / 2.0)
    logP = Z - (Z.max(axis=1, keepdims=True) + np.log(np.exp(Z - Z.max(axis=1, keepdims=True)).sum(axis=1, keepdims=True)))
    rel = logP >= -theta_log2 * np.log(2.0)
    rel = rel | rel.T
    B = Z.shape[0]
    nt = (B + tile - 1) // tile
    pad = nt * tile - B
    rel = np.pad(rel, ((0, pad), (0, pad)))
    return rel.reshape(nt, tile, nt, tile).any(axis=(1, 3))
