import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
from mae_clip_b200 import _lib
from mae_clip_b200._lib import check, ptr, cur_stream
lib = _lib.lib()
B, D, mode = 32768, 256, 1
g = torch.Generator().manual_seed(0)
I = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
T = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
n = lib.mc_clip_loss_fused_workspace_bytes(B, D, mode)
ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
loss = torch.zeros(1, device="cuda"); dI, dT = torch.zeros_like(I), torch.zeros_like(T)
def run():
    check(lib.mc_clip_loss_fwd_bwd(ptr(I), ptr(T), B, D, 1.0, mode, ptr(loss), ptr(dI), ptr(dT), ptr(ws), n, cur_stream()), "x")
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
print("fused device ms/step", a.elapsed_time(b) / 10)
t0 = time.perf_counter()
for _ in range(10): run()
torch.cuda.synchronize()
print("fused wall ms/step", (time.perf_counter() - t0) * 100)
