"""One launch of each HBM-bound kernel at C5 / C2-like sizes (for ncu): python tools/hbm_run.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mae_clip_b200 as m
N, L, P = 1024, 196, 768
x = torch.randn(N, L, P, device="cuda")
noise = torch.rand(N, L, device="cuda")
imgs = torch.randn(N, 3, 224, 224, device="cuda")
pred = torch.randn(N, L, P, device="cuda", requires_grad=True)
xm, mask, restore = m.random_masking(x, 0.75, noise)
loss = m.masked_mse_loss(pred, imgs, mask)
loss.backward()
R = 8192
p = torch.randn(R, R, device="cuda", requires_grad=True)
t = torch.rand(R, R, device="cuda")
out = m.cross_entropy(p, t).sum()
out.backward()
h = m.ProjectionHead(768).cuda().train()
xx = torch.randn(32768, 768, device="cuda")
h(xx).sum().backward()
torch.cuda.synchronize()
print("ok", loss.item())
