"""autograd Functions over the C ABI (``include/mae_clip_b200.h``).

Each Function only marshals pointers/sizes; all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, cur_stream, lib, ptr, require_cuda, workspace


def _mode(mode) -> int:
    if mode is None:  # package default (config.gemm_mode): the fp32-class tensor-core engine
        from . import config as CFG
        mode = CFG.gemm_mode
    if isinstance(mode, int):
        return mode
    try:
        return _lib.GEMM_MODES[mode]
    except KeyError:
        raise ValueError(f"unknown gemm mode {mode!r}; choose from {sorted(_lib.GEMM_MODES)}")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------
# L5: cross_entropy on materialised (rows, cols) tensors          reference: CLIP.py:46-52
# ------------------------------------------------------------------------------------------------
class _SoftCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, targets):
        require_cuda(preds, targets)
        if preds.dim() != 2 or preds.shape != targets.shape:
            raise ValueError("cross_entropy expects 2-D preds/targets of equal shape")
        if preds.dtype != torch.float32:
            preds = preds.float()
        if targets.dtype != torch.float32:
            targets = targets.float()
        rows, cols = preds.shape
        loss = torch.empty(rows, device=preds.device, dtype=torch.float32)
        lse = torch.empty_like(loss)
        tsum = torch.empty_like(loss)
        if rows == 0:  # empty tensors have no storage to point at
            ctx.save_for_backward(preds, targets, lse, tsum)
            return loss
        with torch.cuda.device(preds.device):
            nws = lib().mc_soft_ce_workspace_bytes(rows, cols)
            ws = workspace(nws, preds.device) if nws else None
            check(lib().mc_soft_ce_fwd(ptr(preds), preds.stride(0), preds.stride(1), ptr(targets),
                                       targets.stride(0), targets.stride(1), rows, cols, ptr(loss),
                                       ptr(lse), ptr(tsum), ptr(ws), nws, cur_stream()), "mc_soft_ce_fwd")
        ctx.save_for_backward(preds, targets, lse, tsum)
        return loss

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, grad):
        preds, targets, lse, tsum = ctx.saved_tensors
        rows, cols = preds.shape
        grad = _f32c(grad)
        need_p, need_t = ctx.needs_input_grad
        if rows == 0:
            return (torch.zeros_like(preds) if need_p else None, torch.zeros_like(targets) if need_t else None)
        # gradients take the memory layout of their inputs so stores stay coalesced for `.T` views
        dp = torch.empty_strided(preds.shape, preds.stride(), device=preds.device,
                                 dtype=torch.float32) if need_p and _dense(preds) else (
            torch.empty_like(preds, memory_format=torch.contiguous_format) if need_p else None)
        dt = torch.empty_strided(targets.shape, targets.stride(), device=preds.device,
                                 dtype=torch.float32) if need_t and _dense(targets) else (
            torch.empty_like(targets, memory_format=torch.contiguous_format) if need_t else None)
        with torch.cuda.device(preds.device):
            check(lib().mc_soft_ce_bwd(
                ptr(preds), preds.stride(0), preds.stride(1), ptr(targets), targets.stride(0),
                targets.stride(1), rows, cols, ptr(lse), ptr(tsum), ptr(grad), ptr(dp),
                dp.stride(0) if dp is not None else 0, dp.stride(1) if dp is not None else 0, ptr(dt),
                dt.stride(0) if dt is not None else 0, dt.stride(1) if dt is not None else 0,
                cur_stream()), "mc_soft_ce_bwd")
        return dp, dt


def _dense(t: torch.Tensor) -> bool:
    """True when the 2-D tensor covers its storage span exactly once (plain or transposed)."""
    r, c = t.shape
    s0, s1 = t.stride()
    return (s1 == 1 and s0 == c) or (s0 == 1 and s1 == r)


def soft_cross_entropy_rows(preds: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    return _SoftCE.apply(preds, targets)


# ------------------------------------------------------------------------------------------------
# L3-L6: contrastive soft-target loss on embeddings               reference: CLIP.py:34-43
# ------------------------------------------------------------------------------------------------
class _ClipLoss(torch.autograd.Function):
    """Forward computes the loss and - when a gradient will be needed - dI, dT in the same fused
    call (three sweeps over tiles, nothing of size BxB kept); backward only scales them."""

    @staticmethod
    def forward(ctx, image_emb, text_emb, temperature, mode, grad_mode=True):
        require_cuda(image_emb, text_emb)
        if image_emb.dim() != 2 or image_emb.shape != text_emb.shape:
            raise ValueError("image/text embeddings must both be (B, D)")
        I, T = _f32c(image_emb), _f32c(text_emb)
        B, D = I.shape
        # needs_input_grad mirrors requires_grad even under torch.no_grad(): eval / inference must not pay for the
        # gradient sweep (main.py:115 runs valid_epoch under no_grad).  Grad mode is always off INSIDE forward, so the
        # caller samples it (`grad_mode`).
        need = grad_mode and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        loss = torch.empty((), device=I.device, dtype=torch.float32)
        dI = torch.empty_like(I) if need else None
        dT = torch.empty_like(T) if need else None
        with torch.cuda.device(I.device):
            nbytes = lib().mc_clip_loss_fused_workspace_bytes(B, D, mode)
            ws = workspace(nbytes, I.device)
            check(lib().mc_clip_loss_fwd_bwd(ptr(I), ptr(T), B, D, float(temperature), mode, ptr(loss),
                                             ptr(dI), ptr(dT), ptr(ws), ws.numel(), cur_stream()),
                  "mc_clip_loss_fwd_bwd")
        if need:
            ctx.save_for_backward(dI, dT)
        ctx.in_dtypes = (image_emb.dtype, text_emb.dtype)
        return loss

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, grad_loss):
        dI, dT = ctx.saved_tensors
        gi = (dI * grad_loss).to(ctx.in_dtypes[0]) if ctx.needs_input_grad[0] else None
        gt = (dT * grad_loss).to(ctx.in_dtypes[1]) if ctx.needs_input_grad[1] else None
        return gi, gt, None, None, None


def clip_contrastive_loss(image_emb, text_emb, temperature: float = 1.0, mode=None):
    """Scalar soft-target bidirectional CE of ``CLIPModel.forward`` from (B, D) embeddings."""
    return _ClipLoss.apply(image_emb, text_emb, float(temperature), _mode(mode), torch.is_grad_enabled())


# ------------------------------------------------------------------------------------------------
# L1-L2: ProjectionHead                                           reference: modules.py:55-76
# ------------------------------------------------------------------------------------------------
class _ProjHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_proj, b_proj, w_fc, b_fc, gamma, beta, keep_mask, p_drop, eps, mode, grad_mode=True,
                scale_state=None):
        require_cuda(x, w_proj, b_proj, w_fc, b_fc, gamma, beta, keep_mask)
        lead = x.shape[:-1]
        x2 = _f32c(x.reshape(-1, x.shape[-1]))
        B, E = x2.shape
        P = w_proj.shape[0]
        wp, bp, wf, bf, g, bt = (_f32c(t) for t in (w_proj, b_proj, w_fc, b_fc, gamma, beta))
        if keep_mask is not None:
            keep_mask = keep_mask.reshape(B, P).to(torch.uint8).contiguous()
        need = grad_mode and any(ctx.needs_input_grad[:7])   # no backward state under no_grad (sampled by the caller)
        dev = x2.device
        out = torch.empty(B, P, device=dev, dtype=torch.float32)
        projected = torch.empty(B, P, device=dev, dtype=torch.float32)
        if need:
            hidden = torch.empty_like(projected)
            z = torch.empty_like(projected)
            mean = torch.empty(B, device=dev, dtype=torch.float32)
            rstd = torch.empty_like(mean)
        else:
            hidden = z = mean = rstd = None
        # max|x|, max|hidden| bit patterns (+ two internal words): read by the backward, and - through `scale_state`, a
        # dict the calling module keeps - by the NEXT forward of the same head as the source of its operand scale
        fwd_amax = torch.empty(8, device=dev, dtype=torch.float32) if (need or scale_state is not None) else None  # the C side initialises what it reads
        prev = scale_state.get("amax") if scale_state is not None else None
        if prev is not None and (prev.device != dev or prev.numel() != 8):
            prev = None
        with torch.cuda.device(dev):
            ws = workspace(lib().mc_proj_head_workspace_bytes(B, E, P, mode), dev)
            check(lib().mc_proj_head_fwd(ptr(x2), B, E, P, ptr(wp), ptr(bp), ptr(wf), ptr(bf), ptr(g),
                                         ptr(bt), ptr(keep_mask), float(p_drop), float(eps), mode,
                                         ptr(projected), ptr(hidden), ptr(z), ptr(mean), ptr(rstd),
                                         ptr(out), ptr(fwd_amax), ptr(prev), ptr(ws), ws.numel(), cur_stream()),
                  "mc_proj_head_fwd")
        if scale_state is not None:
            scale_state["amax"] = fwd_amax
        if need:
            ctx.save_for_backward(x2, wp, wf, g, keep_mask, projected, hidden, z, mean, rstd, fwd_amax)
        ctx.cfg = (B, E, P, float(p_drop), mode, lead, x.dtype)
        return out.reshape(*lead, P)

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, grad_out):
        x2, wp, wf, g, keep_mask, projected, hidden, z, mean, rstd, fwd_amax = ctx.saved_tensors
        B, E, P, p_drop, mode, lead, x_dtype = ctx.cfg
        go = _f32c(grad_out.reshape(B, P))
        dev = x2.device
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dwp = torch.empty_like(wp)
        dbp = torch.empty(P, device=dev, dtype=torch.float32)
        dwf = torch.empty_like(wf)
        dbf = torch.empty_like(dbp)
        dg = torch.empty_like(dbp)
        dbt = torch.empty_like(dbp)
        with torch.cuda.device(dev):
            ws = workspace(lib().mc_proj_head_workspace_bytes(B, E, P, mode), dev)
            check(lib().mc_proj_head_bwd(ptr(go), ptr(x2), B, E, P, ptr(wp), ptr(wf), ptr(g),
                                         ptr(keep_mask), p_drop, mode, ptr(projected), ptr(hidden),
                                         ptr(z), ptr(mean), ptr(rstd), ptr(dx), ptr(dwp), ptr(dbp),
                                         ptr(dwf), ptr(dbf), ptr(dg), ptr(dbt), ptr(fwd_amax), ptr(ws), ws.numel(),
                                         cur_stream()), "mc_proj_head_bwd")
        if dx is not None:
            dx = dx.reshape(*lead, E).to(x_dtype)
        return dx, dwp, dbp, dwf, dbf, dg, dbt, None, None, None, None, None, None


def projection_head(x, w_proj, b_proj, w_fc, b_fc, ln_weight, ln_bias, keep_mask=None,
                    p_drop: float = 0.1, eps: float = 1e-5, mode=None, scale_state=None):
    """Fused ProjectionHead forward; ``keep_mask`` (0/1, shape of the output) = training mode.
    ``scale_state``: an optional dict the caller keeps between calls on the same head; the forward leaves the
    max-magnitude words of its input there and takes the operand scale of the next call from them (see
    ``mc_proj_head_fwd`` in the header: saves a pass over ``x``; results do not depend on it)."""
    return _ProjHead.apply(x, w_proj, b_proj, w_fc, b_fc, ln_weight, ln_bias, keep_mask,
                           float(p_drop), float(eps), _mode(mode), torch.is_grad_enabled(), scale_state)


# ------------------------------------------------------------------------------------------------
# M1: random masking                              (not in the reference; oracle/mae_ref.py is spec)
# ------------------------------------------------------------------------------------------------
class _RandomMasking(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, noise, len_keep):
        require_cuda(x, noise)
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise TypeError(f"random_masking: unsupported dtype {x.dtype}")
        x = x.contiguous()
        noise = _f32c(noise)
        N, L, Dm = x.shape
        dev = x.device
        x_masked = torch.empty(N, len_keep, Dm, device=dev, dtype=x.dtype)
        mask = torch.empty(N, L, device=dev, dtype=torch.float32)
        ids_restore = torch.empty(N, L, device=dev, dtype=torch.int64)
        ids_keep = torch.empty(N, len_keep, device=dev, dtype=torch.int64)
        with torch.cuda.device(dev):
            if len_keep > 0 and N > 0:
                check(lib().mc_random_masking(ptr(x), x.element_size(), ptr(noise), N, L, Dm, len_keep,
                                              ptr(x_masked), ptr(mask), ptr(ids_restore), ptr(ids_keep),
                                              cur_stream()), "mc_random_masking")
            elif N > 0:  # nothing kept: indices and mask only (empty tensors have no storage)
                check(lib().mc_random_masking(None, x.element_size(), ptr(noise), N, L, Dm, 0, None,
                                              ptr(mask), ptr(ids_restore), None, cur_stream()),
                      "mc_random_masking")
        ctx.save_for_backward(mask, ids_restore)
        ctx.cfg = (N, L, Dm, len_keep)
        ctx.mark_non_differentiable(mask, ids_restore, ids_keep)
        return x_masked, mask, ids_restore, ids_keep

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, g_masked, _gm, _gr, _gk):
        mask, ids_restore = ctx.saved_tensors
        N, L, Dm, len_keep = ctx.cfg
        if len_keep == 0 or N == 0:   # nothing was kept (mask_ratio 1.0): no token receives a gradient
            return torch.zeros(N, L, Dm, device=g_masked.device, dtype=g_masked.dtype), None, None
        g = g_masked.contiguous()
        gx = torch.empty(N, L, Dm, device=g.device, dtype=g.dtype)
        with torch.cuda.device(g.device):
            check(lib().mc_random_masking_bwd(ptr(g), g.element_size(), ptr(mask), ptr(ids_restore), N,
                                              L, Dm, len_keep, ptr(gx), cur_stream()),
                  "mc_random_masking_bwd")
        return gx, None, None


# ------------------------------------------------------------------------------------------------
# M2-M3: fused patchify + norm-pix target + masked MSE
# ------------------------------------------------------------------------------------------------
class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, imgs, mask, patch, norm_pix):
        require_cuda(pred, imgs, mask)
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        pred = pred.contiguous()
        imgs = _f32c(imgs)
        mask = _f32c(mask)
        N, C, H, W = imgs.shape
        if C != 3:
            raise ValueError("masked_mse expects (N, 3, H, W) images")
        L = (H // patch) * (W // patch)
        if pred.shape != (N, L, patch * patch * 3) or mask.shape != (N, L):
            raise ValueError(f"masked_mse: pred {tuple(pred.shape)} / mask {tuple(mask.shape)} do not "
                             f"match images {tuple(imgs.shape)} at patch {patch}")
        dev = pred.device
        loss = torch.empty((), device=dev, dtype=torch.float32)
        msum = torch.empty((), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            ws = workspace(lib().mc_masked_mse_workspace_bytes(N, L), dev)
            check(lib().mc_masked_mse_fwd(ptr(pred), pred.element_size(), ptr(imgs), ptr(mask), N, H, W,
                                          patch, int(norm_pix), ptr(loss), ptr(msum), ptr(ws),
                                          ws.numel(), cur_stream()), "mc_masked_mse_fwd")
        ctx.save_for_backward(pred, imgs, mask, msum)
        ctx.cfg = (N, H, W, patch, int(norm_pix))
        return loss

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, grad_loss):
        pred, imgs, mask, msum = ctx.saved_tensors
        N, H, W, patch, norm_pix = ctx.cfg
        gl = _f32c(grad_loss)
        dpred = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            check(lib().mc_masked_mse_bwd(ptr(pred), pred.element_size(), ptr(imgs), ptr(mask), N, H, W,
                                          patch, norm_pix, ptr(msum), ptr(gl), ptr(dpred),
                                          cur_stream()), "mc_masked_mse_bwd")
        return dpred, None, None, None, None
