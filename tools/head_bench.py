"""ProjectionHead fwd+bwd timing, fp32 SIMT vs tcgen05: python tools/head_bench.py [B] [tc2048]
(`tc2048`: only the image head on the tcgen05 path - the A/B runs of tools/head_ares_ab.sh)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mae_clip_b200 as m
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ONLY = len(sys.argv) > 2 and sys.argv[2] == "tc2048"
for E in ((2048,) if ONLY else (2048, 768)):
    x = torch.randn(B, E, device="cuda", requires_grad=(E == 2048))
    keep = (torch.rand(B, 256, device="cuda") > 0.1).to(torch.uint8)
    go = torch.randn(B, 256, device="cuda")
    for mode in (("tc_f16x3",) if ONLY else ("simt_fp32", "tc_f16x3")):
        h = m.ProjectionHead(E, gemm_mode=mode).cuda().train()
        def step():
            for p in h.parameters(): p.grad = None
            if x.grad is not None: x.grad = None
            h(x, keep_mask=keep).backward(go)
        for _ in range(3): step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): step()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        flops = 2.0 * B * 256 * (E + 256) * 3 - (0 if E == 2048 else 2.0 * B * E * 256)
        print(f"B={B} E={E} {mode}: {ms:.3f} ms fwd+bwd  ({flops / ms / 1e9:.1f} TFLOP/s algorithmic)", flush=True)
