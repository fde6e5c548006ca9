"""One launch of each cross_entropy kernel on 8192 x 8192 (for ncu): python tools/ce_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mae_clip_b200._lib import check, cur_stream, lib, ptr
L_ = lib()
R = 8192
p = torch.randn(R, R, device="cuda"); t = torch.rand(R, R, device="cuda")
lr, lse, ts = (torch.empty(R, device="cuda") for _ in range(3))
g = torch.ones(R, device="cuda"); dp, dt = torch.empty_like(p), torch.empty_like(t)
nws = L_.mc_soft_ce_workspace_bytes(R, R); ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
for _ in range(2):
    check(L_.mc_soft_ce_fwd(ptr(p), R, 1, ptr(t), R, 1, R, R, ptr(lr), ptr(lse), ptr(ts), ptr(ws), nws, cur_stream()))
    check(L_.mc_soft_ce_fwd(ptr(p), 1, R, ptr(t), 1, R, R, R, ptr(lr), ptr(lse), ptr(ts), ptr(ws), nws, cur_stream()))
    check(L_.mc_soft_ce_bwd(ptr(p), 1, R, ptr(t), 1, R, R, R, ptr(lse), ptr(ts), ptr(g), ptr(dp), 1, R, ptr(dt), 1, R, cur_stream()))
torch.cuda.synchronize(); print("ok")
