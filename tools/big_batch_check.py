import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
import mae_clip_b200 as m
from bench import make_shard
for B in (49152 + 128, 65536, 131072):
    I = torch.cat([make_shard(4096, 100 + i) for i in range((B + 4095) // 4096)])[:B].cuda()
    T = torch.cat([make_shard(4096, 900 + i) for i in range((B + 4095) // 4096)])[:B].cuda()
    res = {}
    for mode in ("tc_f16x3", "simt_fp32") if B <= 65536 else ("tc_f16x3", "tc_f16"):
        Ic, Tc = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
        torch.cuda.synchronize(); t0 = time.time()
        try:
            loss = m.clip_contrastive_loss(Ic, Tc, 1.0, mode=mode)
            loss.backward()
            torch.cuda.synchronize()
            res[mode] = (loss.item(), Ic.grad.clone(), Tc.grad.clone(), time.time() - t0)
        except Exception as e:
            print(B, mode, "FAILED", str(e)[:200]); res[mode] = None
    ks = [k for k in res if res[k]]
    if len(ks) == 2:
        a, b = res[ks[0]], res[ks[1]]
        rel = lambda x, y: ((x - y).double().norm() / y.double().norm()).item()
        print(f"B={B}: loss {ks[0]} {a[0]:.6f} vs {ks[1]} {b[0]:.6f}; dI rel {rel(a[1], b[1]):.2e} dT rel {rel(a[2], b[2]):.2e}; "
              f"times {a[3]*1e3:.1f} / {b[3]*1e3:.1f} ms; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
