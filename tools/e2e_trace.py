"""Timeline of one mc_clip_loss_fwd_bwd_host call (MAE_CLIP_HOST_TRACE=1: CUDA-event marks after every copy and phase,
printed to stderr by the library) for a few inbound-chunk / outbound-strip settings.
Usage (GPU box): python tools/e2e_trace.py [B] 2> trace.log"""
import ctypes
import os
import sys

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
COMBOS = [(1, 1, 0), (1, 2, 0.5), (2, 2, 0.5), (2, 2, 0.375), (2, 2, 0.3125), (2, 2, 0.25),
          (2, 2, 0.1875), (2, 3, 0.125), (2, 3, 0.1875), (4, 2, 0.25)]
if len(sys.argv) > 2 and sys.argv[2] == "default":   # the library's defaults plus their neighbours
    COMBOS = [(4, 2, 0), (4, 3, 0), (8, 2, 0)]
for chunks, strips, last in COMBOS:
    os.environ["MAE_CLIP_HOST_CHUNKS"] = str(chunks)
    os.environ["MAE_CLIP_HOST_STRIPS"] = str(strips)
    os.environ["MAE_CLIP_HOST_LAST_STRIP"] = str(last)
    for it in range(4):
        flush.fill_(1)
        torch.cuda.synchronize()
        if it == 3:
            os.environ["MAE_CLIP_HOST_TRACE"] = "1"
            print(f"---- chunks {chunks} strips {strips} last {last}", file=sys.stderr, flush=True)
        _lib.check(lib.mc_clip_loss_fwd_bwd_host(I.data_ptr(), T.data_ptr(), B, D, 1.0, mode, loss.data_ptr(),
                                                 dI.data_ptr(), dT.data_ptr(), ws.data_ptr(), n, st), "host")
        os.environ.pop("MAE_CLIP_HOST_TRACE", None)
