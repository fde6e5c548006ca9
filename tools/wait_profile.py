"""Where do the warp roles of pair_kernel wait?  Runs the fused single-GPU step on the -DMC_WAIT_PROFILE build
(tools/build_wait_profile.sh) phase by phase and prints, per phase, the cycles each role spent in each class of
mbarrier wait as a share of that role's lifetime (averaged over the CTA pairs).
Usage (GPU box): MAE_CLIP_B200_LIB=tools/prof/libmae_clip_b200_prof.so python tools/wait_profile.py [B]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mae_clip_b200 import _lib  # noqa: E402
from mae_clip_b200._lib import check, ptr, cur_stream  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D, mode = 256, 1
lib = _lib.lib()
raw = C.CDLL(_lib.LIB_PATH)
raw.mc_debug_wait_profile.argtypes = [C.c_void_p, C.c_int]
g = torch.Generator().manual_seed(0)
I = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
T = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).cuda()
n = lib.mc_clip_loss_fused_workspace_bytes(B, D, mode)
ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
loss = torch.zeros(1, device="cuda")
dI, dT = torch.zeros_like(I), torch.zeros_like(T)


def prof(reset=True):
    buf = (C.c_ulonglong * 64)()
    assert raw.mc_debug_wait_profile(buf, 1 if reset else 0) == 0
    return list(buf)


def run(grad):
    check(lib.mc_clip_loss_fwd_bwd(ptr(I), ptr(T), B, D, 1.0, mode, ptr(loss), ptr(dI) if grad else None,
                                   ptr(dT) if grad else None, ptr(ws), n, cur_stream()), "fused")


for _ in range(2):
    run(True)
prof()
run(False)
fwd = prof()
run(True)
both = prof()
bwd = [b - f for b, f in zip(both, fwd)]
names = {"issuer": (6, {0: "A rows", 1: "tile buffer free", 2: "ring slot full", 3: "weights", 4: "X^T tile", 5: "accumulators free"}),
         "producer": (10, {8: "job done", 9: "ring slot free"}),
         "epilogue h0": (20, {16: "constants", 17: "tile full", 18: "grad MMAs done", 19: "accumulators full"}),
         "epilogue h1": (28, {24: "constants", 25: "tile full", 26: "grad MMAs done", 27: "accumulators full"})}
out = {}
for label, v in (("forward sweeps (stats probe + exact Z + row loss)", fwd), ("gradient row sweep (kBwdW)", bwd)):
    out[label] = {}
    for role, (tot, tags) in names.items():
        if v[tot] == 0:
            continue
        out[label][role] = {"lifetime_Mcycles_sum_over_pairs": round(v[tot] / 1e6, 2),
                            **{nm: round(v[t] / v[tot], 3) for t, nm in tags.items()}}
print(json.dumps(out, indent=1))
