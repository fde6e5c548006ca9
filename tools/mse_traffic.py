"""Masked MSE at N = 1024, ratio 0.75: time and (under ncu) DRAM bytes against the algorithmic count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mae_clip_b200 as m
N, L, P = 1024, 196, 768
x = torch.randn(N, L, P, device="cuda"); noise = torch.rand(N, L, device="cuda")
imgs = torch.randn(N, 3, 224, 224, device="cuda")
pred = torch.randn(N, L, P, device="cuda").requires_grad_(True)
_, mask, _ = m.random_masking(x, 0.75, noise)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fwd(): return m.masked_mse_loss(pred.detach(), imgs, mask)
def fb():
    pred.grad = None
    m.masked_mse_loss(pred, imgs, mask).backward()
for name, fn in (("fwd", fwd), ("fwd+bwd", fb)):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print(name, "ms", ts[len(ts) // 2], "algorithmic MB fwd", 0.75 * N * L * P * 8 / 1e6)
