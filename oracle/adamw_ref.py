"""Oracle for the optimiser step of the training driver ("next" row 1, SURVEY.md section 8 f).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

The reference calls ``torch.optim.AdamW(model.parameters(), lr, weight_decay)`` at
``/root/reference/main.py:103-105`` and ``optimizer.step()`` at ``main.py:59``.  The arithmetic
lives in third-party PyTorch (pinned ``torch==2.1.2``, ``requirements.txt:201``; 2.11.0 here), not
under ``/root/reference``; this is a numpy restatement of its published single-tensor algorithm
(decoupled weight decay, bias-corrected moments, ``amsgrad=False``, ``maximize=False``).
``tests/test_oracle_golden.py`` pins it against ``torch.optim.AdamW`` itself on CPU.
"""
from __future__ import annotations

import numpy as np


def adamw_step_ref(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
    """One AdamW step on float32 numpy arrays; returns new (p, m, v).  ``step`` counts from 1."""
    f = np.float32
    p, g, m, v = (np.asarray(a, dtype=f) for a in (p, g, m, v))
    b1, b2 = betas
    p = p * f(1.0 - lr * weight_decay)
    m = m + (g - m) * f(1.0 - b1)
    v = v * f(b2) + (g * g) * f(1.0 - b2)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = np.sqrt(v) / f(bc2 ** 0.5) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)
