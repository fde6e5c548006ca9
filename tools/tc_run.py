"""One fused loss fwd+bwd through the C ABI (for ncu captures): python tools/tc_run.py B mode [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mae_clip_b200 import _lib
from mae_clip_b200._lib import check, ptr, cur_stream

lib = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
D = 256
g = torch.Generator().manual_seed(0)
I = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
T = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
n = lib.mc_clip_loss_fused_workspace_bytes(B, D, mode)
ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
loss = torch.zeros(1, device="cuda")
dI, dT = torch.zeros_like(I), torch.zeros_like(T)
for _ in range(reps):
    check(lib.mc_clip_loss_fwd_bwd(ptr(I), ptr(T), B, D, 1.0, mode, ptr(loss), ptr(dI), ptr(dT), ptr(ws), n,
                                   cur_stream()), "fwd_bwd")
torch.cuda.synchronize()
print("loss", loss.item(), "dI norm", dI.norm().item())
