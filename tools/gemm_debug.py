import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mae_clip_b200 import _lib
from mae_clip_b200._lib import check, ptr, cur_stream
lib = _lib.lib()
for (M, N, K, gelu) in [(128, 128, 64, 0), (128, 128, 128, 0), (256, 256, 256, 1), (24, 256, 160, 1), (1000, 256, 2048, 0), (256, 2048, 1000, 0), (1, 256, 768, 0)]:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda(); B = torch.randn(N, K, generator=g).cuda(); bias = torch.randn(N, generator=g).cuda()
    C = torch.zeros(M, N, device="cuda"); G = torch.zeros(M, N, device="cuda")
    n = lib.mc_tc_gemm_workspace_bytes(M, N, K)
    ws = torch.zeros(n + 512, dtype=torch.uint8, device="cuda")
    check(lib.mc_tc_gemm(ptr(A), ptr(B), M, N, K, ptr(bias), ptr(C), ptr(G) if gelu else None, ptr(ws), n, cur_stream()), "gemm")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T + bias.double()
    err = ((C.double() - ref).norm() / ref.norm()).item()
    print(f"M={M} N={N} K={K} gelu={gelu}: rel err {err:.2e}", (" gelu err %.2e" % ((G.double() - torch.nn.functional.gelu(ref)).norm() / ref.norm()).item()) if gelu else "", flush=True)
