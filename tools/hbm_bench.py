"""HBM-bound kernels through the C ABI with pre-allocated outputs (no allocator / autograd overhead):
achieved GB/s = ALGORITHMIC bytes (SURVEY.md section 8 d) / CUDA-event time (one launch per timing,
median of 15), against the measured copy bandwidth.  Working sets exceed the 126 MB L2 several times.

    python tools/hbm_bench.py [--json out.json] [--flush]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mae_clip_b200 import _lib  # noqa: E402
from mae_clip_b200._lib import check, cur_stream, lib, ptr  # noqa: E402

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L_ = lib()
rows = []


FLUSH = "--flush" in sys.argv  # every working set below exceeds 2x L2, so by default nothing is flushed:
# a 256 MiB WRITE leaves ~126 MB of dirty lines whose write-back is then charged to the timed kernel


def timeit(fn, iters=15, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if FLUSH:
            flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes, **kw):
    r = dict(kernel=name, us=ms * 1e3, algorithmic_MB=nbytes / 1e6, GBps=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / HBM, **kw)
    rows.append(r)
    print(f"{name:44s} {r['us']:9.1f} us  {r['algorithmic_MB']:9.1f} MB  {r['GBps']:7.0f} GB/s  {100 * r['frac_hbm']:5.1f}% of {HBM:.0f}",
          flush=True)


st = cur_stream
N, L, P = 1024, 196, 768
for dtype, es in ((torch.float32, 4), (torch.bfloat16, 2)):
    x = torch.randn(N, L, P, device=dev).to(dtype)
    noise = torch.rand(N, L, device=dev)
    for ratio in (0.5, 0.75, 0.9):
        keep = int(L * (1 - ratio))
        xm = torch.empty(N, keep, P, device=dev, dtype=dtype)
        mask = torch.empty(N, L, device=dev)
        restore = torch.empty(N, L, device=dev, dtype=torch.int64)
        ids_keep = torch.empty(N, keep, device=dev, dtype=torch.int64)
        ms = timeit(lambda: check(L_.mc_random_masking(ptr(x), es, ptr(noise), N, L, P, keep, ptr(xm), ptr(mask), ptr(restore),
                                                       ptr(ids_keep), st())))
        report(f"random_masking {str(dtype)[6:]} r={ratio}", ms, 16 * N * L + 2 * N * keep * P * es, N=N, ratio=ratio)
        if ratio == 0.75:
            gx = torch.empty(N, L, P, device=dev, dtype=dtype)
            ms = timeit(lambda: check(L_.mc_random_masking_bwd(ptr(xm), es, ptr(mask), ptr(restore), N, L, P, keep, ptr(gx), st())))
            report(f"random_masking_bwd {str(dtype)[6:]} r={ratio}", ms, 8 * N * L + N * keep * P * es + N * L * P * es, N=N)
            tok = torch.randn(P, device=dev).to(dtype)
            ms = timeit(lambda: check(L_.mc_restore_tokens(ptr(xm), es, ptr(tok), ptr(restore), N, L, P, keep, ptr(gx), st())))
            report(f"restore_tokens {str(dtype)[6:]} r={ratio}", ms, 8 * N * L + N * keep * P * es + N * L * P * es, N=N)

    imgs = torch.randn(N, 3, 224, 224, device=dev)
    pred = torch.randn(N, L, P, device=dev).to(dtype)
    dpred = torch.empty_like(pred)
    loss = torch.empty((), device=dev)
    msum = torch.empty((), device=dev)
    ws = torch.empty(L_.mc_masked_mse_workspace_bytes(N, L), dtype=torch.uint8, device=dev)
    for ratio in (0.5, 0.75, 0.9):
        keep = int(L * (1 - ratio))
        mask = torch.empty(N, L, device=dev)
        restore = torch.empty(N, L, device=dev, dtype=torch.int64)
        check(L_.mc_random_masking(None, es, ptr(noise), N, L, P, keep, None, ptr(mask), ptr(restore), None, st()))
        r_eff = (L - keep) / L
        ms = timeit(lambda: check(L_.mc_masked_mse_fwd(ptr(pred), es, ptr(imgs), ptr(mask), N, 224, 224, 16, 1, ptr(loss), ptr(msum),
                                                       ptr(ws), ws.numel(), st())))
        bf = r_eff * N * L * P * (es + 4) + 4 * N * L
        report(f"masked_mse_fwd {str(dtype)[6:]} r={ratio}", ms, bf, N=N, ratio=ratio)
        ms = timeit(lambda: check(L_.mc_masked_mse_bwd(ptr(pred), es, ptr(imgs), ptr(mask), N, 224, 224, 16, 1, ptr(msum), None,
                                                       ptr(dpred), st())))
        report(f"masked_mse_bwd {str(dtype)[6:]} r={ratio}", ms, bf + N * L * P * es, N=N, ratio=ratio)
    del x, pred, dpred, imgs

R = 8192
p = torch.randn(R, R, device=dev)
t = torch.rand(R, R, device=dev)
lr, lse, ts = (torch.empty(R, device=dev) for _ in range(3))
g = torch.ones(R, device=dev)
dp, dt = torch.empty_like(p), torch.empty_like(t)
nws = L_.mc_soft_ce_workspace_bytes(R, R)
cws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
ms = timeit(lambda: check(L_.mc_soft_ce_fwd(ptr(p), R, 1, ptr(t), R, 1, R, R, ptr(lr), ptr(lse), ptr(ts), ptr(cws), nws, st())))
report("soft_ce_fwd 8192x8192", ms, 2 * R * R * 4 + 12 * R)
ms = timeit(lambda: check(L_.mc_soft_ce_fwd(ptr(p), 1, R, ptr(t), 1, R, R, R, ptr(lr), ptr(lse), ptr(ts), ptr(cws), nws, st())))
report("soft_ce_fwd 8192x8192 (.T views)", ms, 2 * R * R * 4 + 12 * R)
ms = timeit(lambda: check(L_.mc_soft_ce_bwd(ptr(p), R, 1, ptr(t), R, 1, R, R, ptr(lse), ptr(ts), ptr(g), ptr(dp), R, 1, None, 0, 0, st())))
report("soft_ce_bwd 8192x8192 (dpreds)", ms, 3 * R * R * 4 + 12 * R)
ms = timeit(lambda: check(L_.mc_soft_ce_bwd(ptr(p), R, 1, ptr(t), R, 1, R, R, ptr(lse), ptr(ts), ptr(g), ptr(dp), R, 1, ptr(dt), R, 1, st())))
report("soft_ce_bwd 8192x8192 (dpreds + dtargets)", ms, 4 * R * R * 4 + 12 * R)
ms = timeit(lambda: check(L_.mc_soft_ce_bwd(ptr(p), 1, R, ptr(t), 1, R, R, R, ptr(lse), ptr(ts), ptr(g), ptr(dp), 1, R, ptr(dt), 1, R, st())))
report("soft_ce_bwd 8192x8192 (.T views, both)", ms, 4 * R * R * 4 + 12 * R)

if "--json" in sys.argv:
    json.dump(dict(hbm_peak_gbs=HBM, flush=FLUSH, rows=rows), open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
