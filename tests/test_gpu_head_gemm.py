"""The CTA-pair GEMM kernels the ProjectionHead runs on (csrc/head_tc.cu) through their C-ABI test entry
``mc_head_gemm``, against fp64 matmul: the activation operand is converted fp32 -> fp16 hi / lo INSIDE the kernel
(converter warps -> tensor memory), the weight-gradient form also transposes both operands in the kernel.
Reference arithmetic: modules.py:70-72 (`x Wp^T`, `h Wf^T`) and their autograd (`dp Wp`, `dp^T x`, `dy^T h`)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _call(kind, A, B, M, N, K, bias, want_gelu, passes):
    from mae_clip_b200 import _lib
    from mae_clip_b200._lib import check, cur_stream, ptr
    lib = _lib.lib()
    C = torch.full((M, N), float("nan"), device="cuda")
    G = torch.full((M, N), float("nan"), device="cuda") if want_gelu else None
    n = lib.mc_head_gemm_workspace_bytes(kind, M, N, K)
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    check(lib.mc_head_gemm(kind, ptr(A), ptr(B), M, N, K, ptr(bias), ptr(C), ptr(G), passes, ptr(ws), n, cur_stream()),
          "mc_head_gemm")
    torch.cuda.synchronize()
    return C, G


TOL = {3: 2e-5, 1: 3e-3}


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("M,K", [(256, 64), (128, 128), (300, 160), (2500, 768), (4096, 2048), (20000, 96)])
def test_rows_kernel(M, K, passes):
    g = torch.Generator().manual_seed(M + K)
    A = (torch.randn(M, K, generator=g) * 3).cuda()
    W = torch.randn(256, K, generator=g).cuda()
    bias = torch.randn(256, generator=g).cuda()
    ref = A.double() @ W.double().T + bias.double()
    C, G = _call(0, A, W, M, 256, K, bias, True, passes)
    assert torch.isfinite(C).all() and torch.isfinite(G).all()
    assert rel_err(C, ref) < TOL[passes]
    assert rel_err(G, torch.nn.functional.gelu(ref)) < TOL[passes]
    C2, _ = _call(0, A, W, M, 256, K, None, False, passes)     # plain epilogue, no bias
    assert rel_err(C2, A.double() @ W.double().T) < TOL[passes]


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 160, 96), (2500, 2048, 256), (4096, 768, 256), (300, 1100, 200)])
def test_resident_kernel(M, N, K, passes):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    C, _ = _call(1, A, W, M, N, K, None, False, passes)
    assert torch.isfinite(C).all()
    assert rel_err(C, A.double() @ W.double().T) < TOL[passes]


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("N,K", [(256, 64), (256, 4096), (160, 2500), (768, 5000), (2048, 32768), (2048, 300)])
def test_transposed_operands_kernel(N, K, passes):
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(K, 256, generator=g).cuda()
    B = (torch.randn(K, N, generator=g) * 0.5).cuda()
    C, _ = _call(2, A, B, 256, N, K, None, False, passes)
    assert torch.isfinite(C).all()
    assert rel_err(C, A.double().T @ B.double()) < TOL[passes]
