"""Per-phase device time of the single-GPU loss step through the phase entry points (B = 32768 by default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mae_clip_b200 import _lib
from mae_clip_b200._lib import check, ptr, cur_stream
lib = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D, mode = 256, 1
g = torch.Generator().manual_seed(0)
I = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
T = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
planes = torch.empty(lib.mc_clip_planes_bytes(B, D, mode), dtype=torch.uint8, device="cuda")
nws = lib.mc_clip_loss_workspace_bytes(B, B, D, mode)
ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
r, c, rz, ps = (torch.empty(B, device="cuda") for _ in range(4))
flags = torch.zeros(lib.mc_clip_tile_flags_bytes(B, B, D, mode), dtype=torch.uint8, device="cuda")
check(lib.mc_clip_prepare(ptr(I), ptr(T), B, B, D, 0, mode, ptr(planes), cur_stream()))
def stats(fl):
    check(lib.mc_clip_stats(ptr(I), ptr(T), ptr(planes), B, B, D, 0, 1.0, mode, ptr(r), ptr(c), ptr(rz), ptr(ps), ptr(fl) if fl is not None else None, ptr(ws), nws, cur_stream()))
for fl, name in ((flags, "probe + exact Z on flagged tiles"), (None, "dense (S + 3-pass Z)")):
    for _ in range(2): stats(fl)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): stats(fl)
    b.record(); torch.cuda.synchronize()
    print(os.path.basename(os.environ.get("MAE_CLIP_B200_LIB", "product")), name, "stats ms %.3f" % (a.elapsed_time(b) / 5), "flagged tiles", int(flags.sum().item()) if fl is not None else "-")
