"""A/B of the host-buffer entry's inbound chunk count (MAE_CLIP_HOST_CHUNKS) and outbound strip count
(MAE_CLIP_HOST_STRIPS): wall-clock per call of
mc_clip_loss_fwd_bwd_host at B = 32768, pinned host buffers, L2 flushed between calls.
Usage (GPU box): python tools/e2e_strips.py [B]"""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mae_clip_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D, mode = 256, _lib.GEMM_MODES["tc_f16x3"]
lib = _lib.lib()
g = torch.Generator().manual_seed(0)
I = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).pin_memory()
T = torch.nn.functional.layer_norm(torch.randn(B, D, generator=g), (D,)).pin_memory()
dI, dT, loss = torch.empty_like(I).pin_memory(), torch.empty_like(T).pin_memory(), torch.zeros(1).pin_memory()
n = lib.mc_clip_loss_host_workspace_bytes(B, D, mode)
ws = torch.empty(n, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
out = {}
combos = [(c, s, 0) for c in (1, 2, 4, 8) for s in (1, 2, 4)] + [(4, 3, 0), (4, 6, 0), (16, 2, 0)]
if len(sys.argv) > 2 and sys.argv[2] == "last":     # share of the row blocks the LAST outbound strip takes (0 = equal strips)
    combos = [(4, 2, 0), (4, 2, 0.375), (4, 2, 0.25), (4, 2, 0.125), (4, 3, 0), (4, 3, 0.25), (4, 3, 0.125), (4, 4, 0.125)]
for chunks, strips, last in combos:
    os.environ["MAE_CLIP_HOST_CHUNKS"] = str(chunks)
    os.environ["MAE_CLIP_HOST_STRIPS"] = str(strips)
    os.environ["MAE_CLIP_HOST_LAST_STRIP"] = str(last)
    ts = []
    for it in range(8):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(lib.mc_clip_loss_fwd_bwd_host(I.data_ptr(), T.data_ptr(), B, D, 1.0, mode, loss.data_ptr(),
                                                 dI.data_ptr(), dT.data_ptr(), ws.data_ptr(), n, st), "host")
        ts.append((time.perf_counter() - t0) * 1e3)
    out[f"chunks{chunks}_strips{strips}" + (f"_last{last}" if last else "")] = {"ms_median": sorted(ts[2:])[len(ts[2:]) // 2], "ms_min": min(ts[2:]), "loss": loss.item()}
print(json.dumps(out))
