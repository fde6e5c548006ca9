"""Training-step driver around the hot path ("next" row 1, SURVEY.md section 8 f).

Mirrors ``/root/reference/main.py:51-82`` (``train_epoch`` / ``valid_epoch``) and
``/root/reference/utils.py:1-20`` (``AvgMeter``, ``get_lr``) with the same names, arguments and
return values, minus the two things that stall a B200:

* ``loss.item()`` after every step (``main.py:64,76``) - a host synchronisation per batch.  The
  meter here accumulates ON DEVICE and synchronises only when ``.avg`` / ``.sum`` is read;
* ``torch.optim.AdamW`` 's many small launches - ``AdamW`` below has the same constructor and
  ``state_dict`` layout but its ``step()`` is one multi-tensor kernel launch (``mc_adamw_step``).

``GraphedStep`` captures forward + backward + optimiser step of a static-shape module into one
CUDA graph, for the launch-bound small-batch regime (C2: B = 1024 is ~12 kernels of a few
microseconds each).
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, cur_stream, lib, require_cuda


class AvgMeter:
    """``utils.py:1-16`` with device-side accumulation: ``update`` accepts a python number or a
    0-dim tensor (no ``.item()``); ``avg`` / ``sum`` synchronise lazily when read."""

    def __init__(self, name="Metric"):
        self.name = name
        self.reset()

    def reset(self):
        self._sum_host, self.count = 0.0, 0
        self._sum_dev = None

    def update(self, val, count=1):
        self.count += count
        if isinstance(val, torch.Tensor):
            v = val.detach().to(torch.float64) * count
            self._sum_dev = v if self._sum_dev is None else self._sum_dev + v
        else:
            self._sum_host += val * count

    @property
    def sum(self):
        return self._sum_host + (float(self._sum_dev) if self._sum_dev is not None else 0.0)

    @property
    def avg(self):
        return self.sum / self.count if self.count else 0

    def __repr__(self):
        return f"{self.name}: {self.avg:.4f}"


def get_lr(optimizer):
    """``utils.py:18-20``."""
    for param_group in optimizer.param_groups:
        return param_group["lr"]


class AdamW(torch.optim.Optimizer):
    """Drop-in for ``torch.optim.AdamW(params, lr, weight_decay)`` (``main.py:103-105``): same
    defaults, same per-parameter state (``step``, ``exp_avg``, ``exp_avg_sq``) so ``state_dict()``
    round-trips with torch's; ``step()`` is one ``mc_adamw_step`` launch per 48 tensors.  fp32 CUDA
    parameters only (anything else raises: there is no fallback)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not part of the reference's configuration (main.py:103-105)")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            step = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                require_cuda(p, p.grad)
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise TypeError("mae_clip_b200.AdamW updates dense fp32 parameters only")
                if not p.is_contiguous():
                    raise ValueError("mae_clip_b200.AdamW needs contiguous parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                s = int(st["step"])
                if step is None:
                    step = s
                elif step != s or p.device != ps[0].device:
                    # parameters that joined later (another step count) or live on another device: flush what we have
                    # (one launch covers one device and one bias correction) and continue with theirs
                    self._launch(group, step, ps, gs, ms, vs, grad_scale)
                    ps, gs, ms, vs, step = [], [], [], [], s
                ps.append(p)
                gs.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            if ps:
                self._launch(group, step, ps, gs, ms, vs, grad_scale)
        return loss

    @staticmethod
    def _launch(group, step, ps, gs, ms, vs, grad_scale):
        n = len(ps)
        tab = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
        numel = (C.c_int64 * n)(*[p.numel() for p in ps])
        b1, b2 = group["betas"]
        with torch.cuda.device(ps[0].device):
            check(lib().mc_adamw_step(n, tab(ps), tab(gs), tab(ms), tab(vs), numel, float(group["lr"]), float(b1),
                                      float(b2), float(group["eps"]), float(group["weight_decay"]), step,
                                      None if grad_scale is None else C.c_void_p(grad_scale.data_ptr()), cur_stream()),
                  "mc_adamw_step")


def _to_device(batch, device):
    return {k: v.to(device, non_blocking=True) for k, v in batch.items() if k != "caption"}


def train_epoch(model, train_loader, optimizer, lr_scheduler, step, device=None, progress=None):
    """``main.py:51-67``: same loop, same return value (the loss meter); no per-step host sync.
    ``progress``: optional callable wrapping the loader (e.g. ``tqdm``)."""
    device = device if device is not None else next(model.parameters()).device
    loss_meter = AvgMeter()
    it = progress(train_loader) if progress is not None else train_loader
    for batch in it:
        batch = _to_device(batch, device)
        loss = model(batch)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        if step == "batch":
            lr_scheduler.step()
        loss_meter.update(loss.detach(), batch["image"].size(0))
    return loss_meter


def valid_epoch(model, valid_loader, device=None, progress=None):
    """``main.py:70-82`` (callers wrap it in ``torch.no_grad()`` as ``main.py:115`` does)."""
    device = device if device is not None else next(model.parameters()).device
    loss_meter = AvgMeter()
    it = progress(valid_loader) if progress is not None else valid_loader
    for batch in it:
        batch = _to_device(batch, device)
        loss = model(batch)
        loss_meter.update(loss.detach(), batch["image"].size(0))
    return loss_meter


class GraphedStep:
    """forward + backward (+ optimiser step) of ``loss_fn(*static_inputs)`` captured in ONE CUDA
    graph.  Shapes are fixed at construction; ``__call__(*inputs)`` copies the new inputs into
    the static buffers, replays, and returns the (static) loss tensor.  Gradients live in the
    parameters' ``.grad`` as usual.  With an optimiser, its state tensors are created during the
    warm-up steps (which really update the parameters), exactly like ``torch.cuda.graphs`` usage
    elsewhere; the step counter advances on the host per replay."""

    def __init__(self, loss_fn, example_inputs, params, optimizer=None, warmup=3):
        # inputs that require grad (e.g. embeddings fed to the loss) keep that property on their static copies;
        # their gradients are then read from ``static_in[i].grad`` after a replay
        self.static_in = [t.detach().clone().requires_grad_(t.requires_grad) for t in example_inputs]
        self.params = [p for p in params if p.requires_grad] + [t for t in self.static_in if t.requires_grad]
        self.optimizer = optimizer
        if optimizer is not None and not isinstance(optimizer, AdamW):
            raise TypeError("GraphedStep captures mae_clip_b200.AdamW (its launch carries the step as an argument)")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager(loss_fn)
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(self.graph):
            self.static_loss = loss_fn(*self.static_in)
            self.static_loss.backward()
        self.static_grads = [p.grad for p in self.params]  # written by every replay
        self.loss_fn = loss_fn

    def _eager(self, loss_fn):
        for p in self.params:
            p.grad = None
        loss = loss_fn(*self.static_in)
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss

    def __call__(self, *inputs):
        with torch.no_grad():
            for dst, src in zip(self.static_in, inputs):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.static_grads):
            p.grad = g  # re-attach: callers may have cleared .grad (zero_grad(set_to_none=True)) since the last replay
        if self.optimizer is not None:
            # bias corrections change every step: the optimiser launch stays outside the graph (one launch)
            self.optimizer.step()
        return self.static_loss
