"""Timeline of pair 0's leader CTA inside hg::row_kernel (globaltimer ns), one call per epilogue kind at B = 32768."""
import os
import sys
import torch
sys.path.insert(0, ".")
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["MAE_CLIP_HG_DBG"] = hex(dbg.data_ptr())
import mae_clip_b200 as m  # noqa: E402

B = 32768
names = {0: "entry", 1: "setup done", 2: "mma t0 acc_empty", 3: "mma t0 A ready", 4: "mma t0 B ready", 5: "mma t0 issued",
         6: "mma t1 acc_empty", 7: "mma t1 A ready", 8: "mma t1 B ready", 9: "mma t1 issued", 10: "epi t0 acc_full", 11: "epi t0 done",
         12: "epi t1 acc_full", 13: "epi t1 done", 15: "exit"}
h = m.ProjectionHead(2048, gemm_mode="tc_f16x3").cuda().train()
x = torch.randn(B, 2048, device="cuda", requires_grad=True)
keep = (torch.rand(B, 256, device="cuda") > 0.1).to(torch.uint8)
go = torch.randn(B, 256, device="cuda")
h(x, keep_mask=keep).backward(go)
torch.cuda.synchronize()
from mae_clip_b200 import _lib
from mae_clip_b200._lib import check, cur_stream, ptr
lib = _lib.lib()
for label, K, gelu in (("F1-like x Wp^T K=2048 (bias+GELU)", 2048, True), ("plain K=256", 256, False)):
    A, W = torch.randn(B, K, device="cuda"), torch.randn(256, K, device="cuda")
    bias = torch.randn(256, device="cuda")
    C, G = torch.empty(B, 256, device="cuda"), torch.empty(B, 256, device="cuda")
    n = lib.mc_head_gemm_workspace_bytes(0, B, 256, K)
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dbg.zero_()
        check(lib.mc_head_gemm(0, ptr(A), ptr(W), B, 256, K, ptr(bias), ptr(C), ptr(G) if gelu else None, 3, ptr(ws), n, cur_stream()))
        torch.cuda.synchronize()
    t = dbg.cpu().tolist()
    print(label)
    for i in sorted(names):
        if t[i]:
            print(f"  {names[i]:20s} {(t[i] - t[0]) / 1e3:8.2f} us")
# the head's own F2 / B1 launches: run a full fwd+bwd and read the LAST row-kernel timeline (B1 = GELU backward)
dbg.zero_()
out = h(x, keep_mask=keep)
torch.cuda.synchronize()
t = dbg.cpu().tolist()
print("F2 (LN epilogue), last row kernel of the forward")
for i in sorted(names):
    if t[i]:
        print(f"  {names[i]:20s} {(t[i] - t[0]) / 1e3:8.2f} us")
dbg.zero_()
out.backward(go)
torch.cuda.synchronize()
t = dbg.cpu().tolist()
print("B1 (GELU-backward epilogue), the row kernel of the backward")
for i in sorted(names):
    if t[i]:
        print(f"  {names[i]:20s} {(t[i] - t[0]) / 1e3:8.2f} us")
