"""Oracle for ProjectionHead.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Restates ``/root/reference/modules.py:55-76`` functionally with the dropout
keep-mask as an explicit input (the reference draws it from torch's Philox
stream inside ``nn.Dropout``; parity is defined on "identical inputs and
noise", so tests feed the same mask to both sides).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def proj_head_ref(x, w_proj, b_proj, w_fc, b_fc, ln_w, ln_b, keep_mask=None, p_drop: float = 0.1, eps: float = 1e-5):
    """out = LN(dropout(fc(gelu(proj(x)))) + proj(x)) - ``modules.py:69-76``.

    The residual taps ``projected`` before the GELU (``modules.py:74``); GELU is the
    exact erf form (``nn.GELU()`` default, ``modules.py:64``); LayerNorm eps 1e-5,
    affine (``modules.py:67``).  ``keep_mask`` (B, Dp) of 0/1 reproduces training
    mode (kept values scaled by 1/(1-p)); ``None`` is eval mode.
    """
    projected = F.linear(x, w_proj, b_proj)
    h = F.gelu(projected)
    y = F.linear(h, w_fc, b_fc)
    if keep_mask is not None:
        y = y * keep_mask.to(y.dtype) / (1.0 - p_drop)
    z = y + projected
    return F.layer_norm(z, (z.shape[-1],), ln_w, ln_b, eps)


def proj_head_fwd_bwd_ref(x, params, keep_mask, p_drop, grad_out, need_dx=True, dtype=torch.float32):
    """Forward + autograd backward on CPU; returns (out, dict of grads)."""
    x = x.detach().to(dtype).clone().requires_grad_(need_dx)
    ps = [p.detach().to(dtype).clone().requires_grad_(True) for p in params]
    out = proj_head_ref(x, *ps, keep_mask=keep_mask, p_drop=p_drop)
    out.backward(grad_out.to(dtype))
    names = ["w_proj", "b_proj", "w_fc", "b_fc", "ln_w", "ln_b"]
    grads = {n: p.grad.detach() for n, p in zip(names, ps)}
    grads["x"] = x.grad.detach() if need_dx else None
    return out.detach(), grads
