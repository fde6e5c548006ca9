"""Repeat the fused single-GPU step and the host-buffer entry many times on the same inputs and require bit-identical
loss and gradients every time: a race in one of the mbarrier protocols would show up as a rare mismatch.
Usage (GPU box): python tools/stress_determinism.py [iters]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mae_clip_b200 import _lib  # noqa: E402
from mae_clip_b200._lib import check, ptr, cur_stream  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lib = _lib.lib()
D = 256
bad = 0
for B, mode, scale in [(32768, 1, 1.0), (8192, 1, 1.0), (8192, 2, 1.0), (4224, 1, 1.0), (8192, 1, 0.3), (16384, 1, 1.0)]:
    g = torch.Generator().manual_seed(B + mode)
    I = (torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)) * scale).cuda()
    T = (torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)) * scale).cuda()
    n = lib.mc_clip_loss_fused_workspace_bytes(B, D, mode)
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    loss = torch.zeros(1, device="cuda")
    dI, dT = torch.zeros_like(I), torch.zeros_like(T)
    ref = None
    mism = 0
    for it in range(iters):
        dI.fill_(float("nan")); dT.fill_(float("nan"))
        check(lib.mc_clip_loss_fwd_bwd(ptr(I), ptr(T), B, D, 1.0, mode, ptr(loss), ptr(dI), ptr(dT), ptr(ws), n, cur_stream()), "fused")
        cur = (loss.clone(), dI.clone(), dT.clone())
        if ref is None:
            ref = cur
            assert torch.isfinite(cur[1]).all() and torch.isfinite(cur[2]).all()
        elif not all(torch.equal(a, b) for a, b in zip(ref, cur)):
            mism += 1
    print(f"fused  B={B} mode={mode} scale={scale}: {iters} runs, {mism} mismatches, loss {ref[0].item():.6f}", flush=True)
    bad += mism
    if B in (32768, 16384):
        Ih, Th = I.cpu().pin_memory(), T.cpu().pin_memory()
        oI, oT = torch.empty_like(Ih).pin_memory(), torch.empty_like(Th).pin_memory()
        ol = torch.zeros(1).pin_memory()
        nh = lib.mc_clip_loss_host_workspace_bytes(B, D, mode)
        wsh = torch.zeros(nh, dtype=torch.uint8, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        refh, mism = None, 0
        for it in range(max(iters // 4, 10)):
            oI.fill_(float("nan")); oT.fill_(float("nan"))
            check(lib.mc_clip_loss_fwd_bwd_host(Ih.data_ptr(), Th.data_ptr(), B, D, 1.0, mode, ol.data_ptr(), oI.data_ptr(),
                                                oT.data_ptr(), wsh.data_ptr(), nh, st), "host")
            cur = (ol.clone(), oI.clone(), oT.clone())
            if refh is None:
                refh = cur
                assert torch.isfinite(cur[1]).all() and torch.isfinite(cur[2]).all()
            elif not all(torch.equal(a, b) for a, b in zip(refh, cur)):
                mism += 1
        print(f"host   B={B} mode={mode}: {max(iters // 4, 10)} runs, {mism} mismatches", flush=True)
        bad += mism
print("TOTAL MISMATCHES", bad)
sys.exit(1 if bad else 0)
