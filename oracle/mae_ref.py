"""Oracle for the MAE branch.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

PARITY UNPINNED AGAINST THE REFERENCE: ``/root/reference`` holds no MAE code (no
random_masking, patchify, norm_pix or decoder anywhere - SURVEY.md section 0.2),
so there is no reference file:line to follow.  This file restates the
published MAE formulation (He et al., "Masked Autoencoders Are Scalable Vision
Learners") that BASELINE.json's north_star names: per-sample noise argsort,
keep/restore index gather, normalised-pixel masked-patch MSE.  It is pinned on
the outputs of a published implementation of that formulation, transformers'
ViTMAE (``tests/golden/mae_hf.npz`` from ``tests/golden/make_golden_mae.py``;
checked by ``tests/test_oracle_golden.py``).

Contract fixed here (and matched bit-for-bit by the CUDA kernels):
* ``len_keep = int(L * (1 - mask_ratio))``
* ``ids_shuffle`` = STABLE ascending argsort of the noise (ties -> lower index
  first); ``ids_restore`` = inverse permutation; both int64
* ``mask`` float32, 0 = kept, 1 = removed, in original patch order
* patchify: (N,3,H,W) -> (N, L, p*p*3) with per-patch element order (ph, pw, c)
* norm-pix target: (t - mean) / sqrt(var_unbiased + 1e-6) over the last dim
* loss = sum(mask * mean_p((pred - target)^2)) / sum(mask)
"""
from __future__ import annotations

import numpy as np
import torch


def len_keep_for(L: int, mask_ratio: float) -> int:
    return int(L * (1 - mask_ratio))


def random_masking_ref(x: torch.Tensor, mask_ratio: float, noise: torch.Tensor):
    """x (N, L, D), noise (N, L) in [0,1) -> (x_masked, mask, ids_restore, ids_keep)."""
    N, L, D = x.shape
    keep = len_keep_for(L, mask_ratio)
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, :keep]
    x_masked = torch.gather(x, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))
    mask = torch.ones(N, L, dtype=torch.float32)
    mask[:, :keep] = 0
    mask = torch.gather(mask, 1, ids_restore)
    return x_masked, mask, ids_restore, ids_keep


def random_masking_numpy(noise: np.ndarray, mask_ratio: float):
    """Index-only numpy restatement (independent of torch's sort) - bit-exact check."""
    N, L = noise.shape
    keep = len_keep_for(L, mask_ratio)
    ids_shuffle = np.argsort(noise, axis=1, kind="stable").astype(np.int64)
    ids_restore = np.empty_like(ids_shuffle)
    rows = np.arange(N)[:, None]
    ids_restore[rows, ids_shuffle] = np.arange(L, dtype=np.int64)[None, :]
    mask = (ids_restore >= keep).astype(np.float32)
    return ids_shuffle[:, :keep], ids_restore, mask


def patchify_ref(imgs: torch.Tensor, p: int = 16) -> torch.Tensor:
    """(N, 3, H, W) -> (N, (H/p)*(W/p), p*p*3), element order (ph, pw, c)."""
    N, C, H, W = imgs.shape
    h, w = H // p, W // p
    x = imgs.reshape(N, C, h, p, w, p)
    x = x.permute(0, 2, 4, 3, 5, 1)  # n h w p q c
    return x.reshape(N, h * w, p * p * C)


def unpatchify_ref(x: torch.Tensor, p: int = 16, C: int = 3) -> torch.Tensor:
    N, L, _ = x.shape
    h = w = int(round(L ** 0.5))
    x = x.reshape(N, h, w, p, p, C).permute(0, 5, 1, 3, 2, 4)
    return x.reshape(N, C, h * p, w * p)


def norm_pix_target_ref(imgs: torch.Tensor, p: int = 16, eps: float = 1e-6) -> torch.Tensor:
    t = patchify_ref(imgs, p)
    mean = t.mean(dim=-1, keepdim=True)
    var = t.var(dim=-1, keepdim=True)  # unbiased
    return (t - mean) / (var + eps) ** 0.5


def masked_mse_ref(pred: torch.Tensor, imgs: torch.Tensor, mask: torch.Tensor, p: int = 16,
                   norm_pix: bool = True) -> torch.Tensor:
    """pred (N, L, p*p*3), imgs (N,3,H,W), mask (N, L) -> scalar."""
    target = norm_pix_target_ref(imgs, p) if norm_pix else patchify_ref(imgs, p)
    per_patch = ((pred.float() - target.float()) ** 2).mean(dim=-1)
    return (per_patch * mask).sum() / mask.sum()


def masked_mse_fwd_bwd_ref(pred, imgs, mask, p: int = 16, norm_pix: bool = True, dtype=torch.float32):
    pred = pred.detach().to(dtype).clone().requires_grad_(True)
    loss = masked_mse_ref(pred, imgs.to(dtype), mask.to(dtype), p, norm_pix)
    loss.backward()
    return loss.detach(), pred.grad.detach()


def restore_tokens_ref(x_kept: torch.Tensor, mask_token: torch.Tensor, ids_restore: torch.Tensor):
    """Decoder-side glue (SURVEY.md section 8 f rank 2): append mask tokens, un-shuffle."""
    N, keep, D = x_kept.shape
    L = ids_restore.shape[1]
    fill = mask_token.reshape(1, 1, D).expand(N, L - keep, D)
    full = torch.cat([x_kept, fill], dim=1)
    return torch.gather(full, 1, ids_restore.unsqueeze(-1).expand(-1, -1, D))
