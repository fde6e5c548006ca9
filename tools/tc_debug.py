"""GPU bring-up script for the tcgen05 engine: phase-by-phase comparison with the fp32 SIMT engine
through the C ABI (not a test; run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mae_clip_b200 import _lib
from mae_clip_b200._lib import check, ptr, cur_stream

lib = _lib.lib()


def emb(B, D, seed, scale):
    g = torch.Generator().manual_seed(seed)
    return (torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)) * scale).cuda()


def phases(I, T, tau, mode, b=None, off=0):
    B, D = I.shape
    b = b or B
    dev = I.device
    nb = lib.mc_clip_planes_bytes(B, D, mode)
    planes = torch.zeros(max(nb, 1), dtype=torch.uint8, device=dev)
    ws = torch.zeros(max(lib.mc_clip_loss_workspace_bytes(b, B, D, mode), 1), dtype=torch.uint8, device=dev)
    st = cur_stream()
    check(lib.mc_clip_prepare(ptr(I), ptr(T), B, B, D, 0, mode, ptr(planes), st), "prepare")
    s = torch.zeros(4, b, device=dev)
    check(lib.mc_clip_stats(ptr(I), ptr(T), ptr(planes), b, B, D, off, tau, mode, ptr(s[0]), ptr(s[1]), ptr(s[2]),
                            ptr(s[3]), ptr(ws), ws.numel(), st), "stats")
    torch.cuda.synchronize()
    return planes, ws, s


def full(I, T, tau, mode):
    B, D = I.shape
    dev = I.device
    n = lib.mc_clip_loss_fused_workspace_bytes(B, D, mode)
    ws = torch.zeros(n, dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, device=dev)
    dI, dT = torch.zeros_like(I), torch.zeros_like(T)
    check(lib.mc_clip_loss_fwd_bwd(ptr(I), ptr(T), B, D, tau, mode, ptr(loss), ptr(dI), ptr(dT), ptr(ws), n,
                                   cur_stream()), "fwd_bwd")
    torch.cuda.synchronize()
    return loss, dI, dT


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


which = sys.argv[1] if len(sys.argv) > 1 else "stats"
sizes = [(128, 256, 0.1), (256, 256, 1.0), (1000, 256, 0.25), (4096, 256, 1.0), (384, 128, 0.3)]
for B, D, scale in sizes:
    I, T = emb(B, D, 0, scale), emb(B, D, 1, scale)
    if which == "stats":
        _, _, s0 = phases(I, T, 1.0, 0)
        for mode in (1, 2):
            _, _, s1 = phases(I, T, 1.0, mode)
            print(f"B={B} D={D} scale={scale} mode={mode} stats maxabs diff r/c/rz:",
                  [(s1[k] - s0[k]).abs().max().item() for k in range(4)], "ref range", s0.min().item(), s0.max().item(),
                  flush=True)
            if not torch.isfinite(s1).all() or (s1 - s0).abs().max() > 1e-2:
                print("   first rows tc:", s1[:, :4].tolist(), " simt:", s0[:, :4].tolist())
    else:
        l0, dI0, dT0 = full(I, T, 1.0, 0)
        for mode in (1, 2):
            l1, dI1, dT1 = full(I, T, 1.0, mode)
            print(f"B={B} D={D} scale={scale} mode={mode} loss {l1.item():.6f} vs {l0.item():.6f} "
                  f"rel dI {rel(dI1, dI0):.2e} dT {rel(dT1, dT0):.2e}", flush=True)

if which == "rowloss":
    for B, D, scale in [(256, 256, 1.0), (4096, 256, 1.0)]:
        I, T = emb(B, D, 0, scale), emb(B, D, 1, scale)
        res = {}
        for mode in (0, 1):
            planes, ws, s = phases(I, T, 1.0, mode)
            gq = torch.zeros(2, B, device="cuda"); part = torch.zeros(1, device="cuda")
            check(lib.mc_clip_rowloss(ptr(I), ptr(T), ptr(planes), B, B, D, 0, 1.0, mode, ptr(s[0]), ptr(s[1]), ptr(s[2]), ptr(s[3]),
                                      ptr(gq[0]), ptr(gq[1]), ptr(part), ptr(ws), ws.numel(), cur_stream()), "rowloss")
            torch.cuda.synchronize()
            res[mode] = (s.clone(), gq.clone(), part.clone())
        s0, gq0, p0 = res[0]; s1, gq1, p1 = res[1]
        print(f"B={B}: stats maxdiff {[(s1[k]-s0[k]).abs().max().item() for k in range(3)]}  g rel {rel(gq1[0], gq0[0]):.2e} "
              f"q maxdiff {(gq1[1]-gq0[1]).abs().max().item():.2e} loss {p1.item():.6f} vs {p0.item():.6f}")
        # feed SIMT stats into the TC bwd and vice versa to localise the error
        for mode_b, (ss, gg) in (("tc bwd with simt stats", (s0, gq0)), ("tc bwd with tc stats", (s1, gq1))):
            planes, ws, _ = phases(I, T, 1.0, 1)
            dI = torch.zeros_like(I); dT = torch.zeros_like(T)
            check(lib.mc_clip_bwd(ptr(I), ptr(T), ptr(planes), B, B, D, 0, 1.0, 1, ptr(ss[0]), ptr(ss[1]), ptr(ss[2]),
                                  ptr(gg[0]), ptr(gg[1]), None, ptr(dI), ptr(dT), ptr(ws), ws.numel(), cur_stream()), "bwd")
            torch.cuda.synchronize()
            l0, dI0, dT0 = full(I, T, 1.0, 0)
            print("   ", mode_b, f"rel dI {rel(dI, dI0):.2e} dT {rel(dT, dT0):.2e}")
