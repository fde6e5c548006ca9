"""Blockwise float64 oracle of the contrastive soft-target loss for batches whose B x B matrices
do not fit anywhere (B = 8192 ... 32768: BASELINE config 4).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  It is the closed form of
``oracle.loss_ref.clip_loss_closed_form`` (derived from
``/root/reference/CLIP.py:34-43``) evaluated one strip of rows at a time with plain torch float64
ops, on whatever device the inputs live on - the parity tests run it ON THE GPU as the checker of the
CUDA path at the benchmarked size; ``tests/test_oracle_golden.py`` pins it on the numpy closed form
(and through it on the reference fixtures) on CPU.

Three passes over row strips of ``rows`` samples (each strip holds a few ``rows x B`` fp64 matrices):

  1. S = T_i I^T / tau, Z = (I_i I^T + T_i T^T) tau/2  ->  r = rowLSE(S), rz = rowLSE(Z) of the strip;
     running column (max, sum) of S -> c = colLSE(S); P = exp(Z - rz) -> q = colsum(P)
  2. G = -(2 S - r_i - c_j) / 2B  ->  g_i = sum_j P_ij G_ij,  loss = sum_i g_i
  3. dS strip, the TRANSPOSED strip dS_ji (from S^T-strip = I_i T^T / tau and the symmetry of Z) and
     dZ + dZ^T  ->  dT_i = dS I / tau + (tau/2) dZs T,  dI_i = dS^T-strip T / tau + (tau/2) dZs I
"""
from __future__ import annotations

import torch


def _lse_rows(a: torch.Tensor) -> torch.Tensor:
    m = a.max(dim=1, keepdim=True).values
    return (m + (a - m).exp().sum(dim=1, keepdim=True).log()).squeeze(1)


def clip_loss_blockwise_f64(image_emb: torch.Tensor, text_emb: torch.Tensor, temperature: float = 1.0,
                            rows: int = 1024, grad_loss: float = 1.0, want_grad: bool = True):
    """Returns ``(loss, dI, dT, stats)``: python float, two (B, D) float64 tensors on the input device (``None``
    when ``want_grad`` is False) and the dict of the five length-B statistics the kernels keep."""
    I = image_emb.detach().to(torch.float64)
    T = text_emb.detach().to(torch.float64)
    B, D = I.shape
    tau = float(temperature)
    dev = I.device
    r = torch.empty(B, dtype=torch.float64, device=dev)
    rz = torch.empty_like(r)
    cm = torch.full((B,), -float("inf"), dtype=torch.float64, device=dev)   # running column max of S
    cs = torch.zeros(B, dtype=torch.float64, device=dev)                    # running column sum relative to cm
    q = torch.zeros(B, dtype=torch.float64, device=dev)
    strips = [(i0, min(i0 + rows, B)) for i0 in range(0, B, rows)]
    for i0, i1 in strips:                                                   # ---- pass 1
        S = (T[i0:i1] @ I.T) / tau
        Z = (I[i0:i1] @ I.T + T[i0:i1] @ T.T) * (tau / 2.0)
        r[i0:i1] = _lse_rows(S)
        rz[i0:i1] = _lse_rows(Z)
        m_new = torch.maximum(cm, S.max(dim=0).values)
        cs = cs * (cm - m_new).exp() + (S - m_new).exp().sum(dim=0)
        cm = m_new
        q += (Z - rz[i0:i1, None]).exp().sum(dim=0)
    c = cm + cs.log()
    g = torch.empty(B, dtype=torch.float64, device=dev)
    for i0, i1 in strips:                                                   # ---- pass 2
        S = (T[i0:i1] @ I.T) / tau
        Z = (I[i0:i1] @ I.T + T[i0:i1] @ T.T) * (tau / 2.0)
        P = (Z - rz[i0:i1, None]).exp()
        G = -(2.0 * S - r[i0:i1, None] - c[None, :]) / (2.0 * B)
        g[i0:i1] = (G * P).sum(dim=1)
    loss = g.sum().item()
    stats = {"row_lse_s": r, "col_lse_s": c, "row_lse_z": rz, "row_g": g, "col_sum_p": q}
    if not want_grad:
        return loss, None, None, stats
    dI = torch.empty(B, D, dtype=torch.float64, device=dev)
    dT = torch.empty_like(dI)
    for i0, i1 in strips:                                                   # ---- pass 3
        S = (T[i0:i1] @ I.T) / tau                                          # S_ij,  i in the strip
        St = (I[i0:i1] @ T.T) / tau                                         # S_ji as [i, j]
        Z = (I[i0:i1] @ I.T + T[i0:i1] @ T.T) * (tau / 2.0)                 # Z_ij = Z_ji
        P = (Z - rz[i0:i1, None]).exp()                                     # P_ij
        Pt = (Z - rz[None, :]).exp()                                        # P_ji as [i, j]
        dS = ((S - r[i0:i1, None]).exp() + (S - c[None, :]).exp() * q[None, :] - 2.0 * P) / (2.0 * B)
        dSt = ((St - r[None, :]).exp() + (St - c[i0:i1, None]).exp() * q[i0:i1, None] - 2.0 * Pt) / (2.0 * B)
        G = -(2.0 * S - r[i0:i1, None] - c[None, :]) / (2.0 * B)
        Gt = -(2.0 * St - r[None, :] - c[i0:i1, None]) / (2.0 * B)
        dZs = P * (G - g[i0:i1, None]) + Pt * (Gt - g[None, :])             # dZ_ij + dZ_ji
        dT[i0:i1] = (dS @ I) / tau + (tau / 2.0) * (dZs @ T)
        dI[i0:i1] = (dSt @ T) / tau + (tau / 2.0) * (dZs @ I)
    return loss, dI * grad_loss, dT * grad_loss, stats
