"""One ProjectionHead forward + backward per head on the tcgen05 path, for an ncu launch list:
  ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/head_tc_launches.csv python tools/head_tc_run.py
(image head E = 2048 with dx, then text head E = 768 without; B = 32768; one warm-up step outside the profiled range)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mae_clip_b200 as m  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rt = torch.cuda.cudart()
for E, need_dx in ((2048, True), (768, False)):
    h = m.ProjectionHead(E, gemm_mode="tc_f16x3").cuda().train()
    x = torch.randn(B, E, device="cuda", requires_grad=need_dx)
    keep = (torch.rand(B, 256, device="cuda") > 0.1).to(torch.uint8)
    go = torch.randn(B, 256, device="cuda")
    h(x, keep_mask=keep).backward(go)          # warm-up (attribute set-up, allocator)
    for p in h.parameters():
        p.grad = None
    x.grad = None
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    h(x, keep_mask=keep).backward(go)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    print(f"head E={E}: done", flush=True)
