"""Head fwd+bwd error vs fp64 at B=32768 (and time), for a given env configuration."""
import sys, torch
sys.path.insert(0, ".")
import mae_clip_b200 as m
torch.manual_seed(0)
B = 32768
for E, need_dx in ((2048, True), (768, False)):
    h = m.ProjectionHead(E, gemm_mode="tc_f16x3").cuda().train()
    x = torch.randn(B, E, device="cuda", requires_grad=need_dx)
    keep = (torch.rand(B, 256, device="cuda") > 0.1).to(torch.uint8)
    go = torch.randn(B, 256, device="cuda")
    out = h(x, keep_mask=keep); out.backward(go)
    # fp64 reference on the GPU
    hd = torch.nn.ModuleDict()
    W = {k: v.detach().double().requires_grad_(True) for k, v in h.state_dict().items()}
    xd = x.detach().double().requires_grad_(need_dx)
    p = xd @ W["projection.weight"].T + W["projection.bias"]
    hh = torch.nn.functional.gelu(p)
    y = hh @ W["fc.weight"].T + W["fc.bias"]
    z = y * keep.double() / 0.9 + p
    o = torch.nn.functional.layer_norm(z, (256,), W["layer_norm.weight"], W["layer_norm.bias"], 1e-5)
    o.backward(go.double())
    rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
    print(f"E={E}: out {rel(out, o):.2e}", " ".join(f"{k}:{rel(dict(h.named_parameters())[k].grad, W[k].grad):.2e}" for k in W),
          f"dx {rel(x.grad, xd.grad):.2e}" if need_dx else "")
    def step():
        for q in h.parameters(): q.grad = None
        x.grad = None
        h(x, keep_mask=keep).backward(go)
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): step()
    b.record(); torch.cuda.synchronize()
    print(f"   {a.elapsed_time(b) / 10:.3f} ms fwd+bwd")
