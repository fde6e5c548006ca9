import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mae_clip_b200 as m
torch.manual_seed(0)
B, E = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (24, 160)
h0 = m.ProjectionHead(E, gemm_mode="simt_fp32").cuda().eval()
h1 = m.ProjectionHead(E, gemm_mode="tc_f16x3").cuda().eval()
h1.load_state_dict(h0.state_dict())
x = torch.randn(B, E, device="cuda")
with torch.no_grad():
    a = h0(x); torch.cuda.synchronize(); print("simt ok")
    b = h1(x); torch.cuda.synchronize(); print("tc ok")
print("rel err", ((a - b).norm() / a.norm()).item())
