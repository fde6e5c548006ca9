"""One call of each head GEMM kernel at the benchmark shapes (B = 32768) for an ncu launch list / timing."""
import sys
import time
import torch
sys.path.insert(0, ".")
from mae_clip_b200 import _lib
from mae_clip_b200._lib import check, cur_stream, ptr

lib = _lib.lib()
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cases = [("x Wp^T  (rows, K=2048)", 0, Bn, 256, 2048), ("h Wf^T  (rows, K=256)", 0, Bn, 256, 256),
         ("x Wp^T  (rows, K=768)", 0, Bn, 256, 768),
         ("dx = dp Wp (resident, N=2048)", 1, Bn, 2048, 256), ("dh = dy Wf (resident, N=256)", 1, Bn, 256, 256),
         ("dWp = dp^T x (tt, N=2048)", 2, 256, 2048, Bn), ("dWp = dp^T x (tt, N=768)", 2, 256, 768, Bn),
         ("dWf = dy^T h (tt, N=256)", 2, 256, 256, Bn)]
for name, kind, M, N, K in cases:
    if kind == 2:
        A, B = torch.randn(K, 256, device="cuda"), torch.randn(K, N, device="cuda")
    else:
        A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    C = torch.empty(M, N, device="cuda")
    G = torch.empty(M, N, device="cuda") if kind == 0 else None
    n = lib.mc_head_gemm_workspace_bytes(kind, M, N, K)
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    def run():
        check(lib.mc_head_gemm(kind, ptr(A), ptr(B), M, N, K, ptr(bias), ptr(C), ptr(G), 3, ptr(ws), n, cur_stream()))
    run(); run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        run()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / iters * 1e3:.1f} us per call (incl. amax + weight staging)")
