"""MAE branch of the hot path: per-sample random masking and normalised-pixel masked MSE.

The reference tree has no MAE code; the API follows the published MAE
formulation that BASELINE.json's north_star names and ``oracle/mae_ref.py`` restates.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import functional as F_b200
from ._lib import check, cur_stream, lib, ptr, require_cuda


def len_keep_for(L: int, mask_ratio: float) -> int:
    return int(L * (1 - mask_ratio))


def random_masking(x: torch.Tensor, mask_ratio: float, noise: torch.Tensor | None = None):
    """x (N, L, D) -> (x_masked (N, len_keep, D), mask (N, L) f32 0=keep/1=removed,
    ids_restore (N, L) int64).  ``noise`` (N, L) in [0,1) is drawn with torch's generator when not
    given; the argsort is stable ascending, so results are reproducible bit for bit."""
    require_cuda(x, noise)
    N, L, _ = x.shape
    if noise is None:
        noise = torch.rand(N, L, device=x.device)
    x_masked, mask, ids_restore, _ids_keep = F_b200._RandomMasking.apply(x, noise,
                                                                         len_keep_for(L, mask_ratio))
    return x_masked, mask, ids_restore


def random_masking_with_ids(x, mask_ratio, noise=None):
    require_cuda(x, noise)
    N, L, _ = x.shape
    if noise is None:
        noise = torch.rand(N, L, device=x.device)
    return F_b200._RandomMasking.apply(x, noise, len_keep_for(L, mask_ratio))


def masked_mse_loss(pred, imgs, mask, patch_size: int = 16, norm_pix_loss: bool = True):
    """sum(mask * mean_e (pred - target)^2) / sum(mask); target = patchify(imgs), normalised per
    patch when ``norm_pix_loss``.  The target is never written to memory."""
    return F_b200._MaskedMSE.apply(pred, imgs, mask, int(patch_size), bool(norm_pix_loss))


def patchify(imgs: torch.Tensor, patch_size: int = 16, norm_pix: bool = False) -> torch.Tensor:
    """(N, 3, H, W) -> (N, L, p*p*3), element order (ph, pw, c).  No autograd (a data op)."""
    require_cuda(imgs)
    imgs = imgs.float().contiguous()
    N, C, H, W = imgs.shape
    if C != 3:
        raise ValueError("patchify expects (N, 3, H, W)")
    L = (H // patch_size) * (W // patch_size)
    out = torch.empty(N, L, patch_size * patch_size * 3, device=imgs.device, dtype=torch.float32)
    with torch.cuda.device(imgs.device):
        check(lib().mc_patchify(ptr(imgs), N, H, W, patch_size, int(norm_pix), ptr(out), cur_stream()),
              "mc_patchify")
    return out


class _RestoreTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_kept, mask_token, ids_restore):
        require_cuda(x_kept, mask_token, ids_restore)
        if x_kept.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise TypeError(f"restore_tokens: unsupported dtype {x_kept.dtype}")
        x_kept = x_kept.contiguous()
        token = mask_token.to(x_kept.dtype).reshape(-1).contiguous()
        ids_restore = ids_restore.contiguous()
        N, keep, D = x_kept.shape
        L = ids_restore.shape[1]
        out = torch.empty(N, L, D, device=x_kept.device, dtype=x_kept.dtype)
        with torch.cuda.device(x_kept.device):
            check(lib().mc_restore_tokens(ptr(x_kept), x_kept.element_size(), ptr(token), ptr(ids_restore), N, L, D,
                                          keep, ptr(out), cur_stream()), "mc_restore_tokens")
        ctx.save_for_backward(ids_restore)
        ctx.cfg = (N, L, D, keep, mask_token.shape, mask_token.dtype)
        return out

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, grad_out):
        ids_restore, = ctx.saved_tensors
        N, L, D, keep, tok_shape, tok_dtype = ctx.cfg
        g = grad_out.contiguous()
        dx = torch.empty(N, keep, D, device=g.device, dtype=g.dtype)
        dtok = torch.empty(D, device=g.device, dtype=g.dtype)
        with torch.cuda.device(g.device):
            from ._lib import workspace
            ws = workspace(lib().mc_restore_tokens_bwd_workspace_bytes(N, L, D), g.device)
            check(lib().mc_restore_tokens_bwd(ptr(g), g.element_size(), int(g.dtype == torch.bfloat16), ptr(ids_restore), N,
                                              L, D, keep, ptr(dx), ptr(dtok), ptr(ws), ws.numel(), cur_stream()),
                  "mc_restore_tokens_bwd")
        return dx, dtok.reshape(tok_shape).to(tok_dtype), None


def restore_tokens(x_kept: torch.Tensor, mask_token: torch.Tensor, ids_restore: torch.Tensor):
    """Decoder-side glue (SURVEY.md section 8 f rank 2): (N, len_keep, D) kept tokens + a (D,) mask
    token -> (N, L, D) in original patch order.  Differentiable in the kept tokens and the mask token."""
    return _RestoreTokens.apply(x_kept, mask_token, ids_restore)
