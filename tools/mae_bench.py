"""C5 sweep: MAE random masking + normalised-pixel masked MSE, achieved HBM GB/s against the
algorithmic bytes of SURVEY.md section 8(d).  python tools/mae_bench.py [--json out]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mae_clip_b200 as m

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
L, P = 196, 768
for dtype, s in ((torch.float32, 4), (torch.bfloat16, 2)):
    for N in (64, 256, 1024):
        x = torch.randn(N, L, P, device="cuda").to(dtype)
        noise = torch.rand(N, L, device="cuda")
        imgs = torch.randn(N, 3, 224, 224, device="cuda")
        pred = torch.randn(N, L, P, device="cuda").to(dtype).requires_grad_(True)
        for ratio in (0.5, 0.75, 0.9):
            keep = int(L * (1 - ratio))
            t1 = timeit(lambda: m.random_masking(x, ratio, noise))
            b1 = 16 * N * L + 2 * N * keep * P * s
            _, mask, _ = m.random_masking(x, ratio, noise)
            r_eff = (L - keep) / L
            t2 = timeit(lambda: m.masked_mse_loss(pred.detach(), imgs, mask))
            b2 = r_eff * N * L * P * (s + 4) + 4 * N * L

            def fb():
                pred.grad = None
                m.masked_mse_loss(pred, imgs, mask).backward()
            t3 = timeit(fb)
            b3 = 2 * b2 + N * L * P * s
            rows.append(dict(dtype=str(dtype).split(".")[-1], N=N, ratio=ratio,
                             masking_us=t1 * 1e3, masking_GBps=b1 / t1 / 1e6, masking_frac=b1 / t1 / 1e6 / HBM,
                             mse_fwd_us=t2 * 1e3, mse_fwd_GBps=b2 / t2 / 1e6, mse_fwd_frac=b2 / t2 / 1e6 / HBM,
                             mse_fwd_bwd_us=t3 * 1e3, mse_fwd_bwd_GBps=b3 / t3 / 1e6, mse_fwd_bwd_frac=b3 / t3 / 1e6 / HBM))
            r = rows[-1]
            print(f"{r['dtype']:8s} N={N:5d} r={ratio:.2f}  masking {r['masking_us']:8.1f} us {r['masking_frac']*100:5.1f}%  "
                  f"mse fwd {r['mse_fwd_us']:8.1f} us {r['mse_fwd_frac']*100:5.1f}%  fwd+bwd {r['mse_fwd_bwd_us']:8.1f} us {r['mse_fwd_bwd_frac']*100:5.1f}%",
                  flush=True)
if len(sys.argv) > 2 and sys.argv[1] == "--json":
    json.dump(dict(hbm_peak_gbs=HBM, rows=rows), open(sys.argv[2], "w"), indent=1)
