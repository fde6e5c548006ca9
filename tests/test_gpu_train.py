"""Training-step driver ("next" row 1, SURVEY.md section 8 f): multi-tensor AdamW through the C ABI
against the oracle (numpy restatement, itself pinned to torch.optim.AdamW on CPU by
tests/test_oracle_golden.py) and against torch's own CUDA AdamW; train_epoch / valid_epoch against the
reference loop of main.py:51-82 with the per-step ``.item()``; CUDA-graph replay against eager."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import adamw_ref, loss_ref

pytestmark = pytest.mark.gpu

SHAPES = [(256, 2048), (256,), (256, 256), (7,), (3, 5, 11), (1,), (1025,)]


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in SHAPES]


@pytest.mark.parametrize("wd,lr", [(1e-3, 1e-3), (0.0, 3e-2), (0.1, 1e-4)])
def test_adamw_vs_oracle_and_torch(wd, lr):
    from mae_clip_b200.train import AdamW
    p0 = _make(0)
    ours = [p.clone().cuda().requires_grad_(True) for p in p0]
    theirs = [p.clone().cuda().requires_grad_(True) for p in p0]
    ref = [(p.numpy().copy(), np.zeros(p.shape, np.float32), np.zeros(p.shape, np.float32)) for p in p0]
    opt = AdamW(ours, lr=lr, weight_decay=wd)
    topt = torch.optim.AdamW(theirs, lr=lr, weight_decay=wd)
    for step in range(1, 6):
        grads = _make(100 + step)
        for p, q, g in zip(ours, theirs, grads):
            p.grad = g.cuda()
            q.grad = g.cuda()
        opt.step()
        topt.step()
        ref = [adamw_ref.adamw_step_ref(p, g.numpy(), m, v, step, lr=lr, weight_decay=wd) for (p, m, v), g in zip(ref, grads)]
    for p, q, (rp, rm, rv) in zip(ours, theirs, ref):
        assert rel_err(p.detach(), rp) < 2e-6
        assert rel_err(p.detach(), q.detach()) < 2e-6
        st = opt.state[p]
        assert rel_err(st["exp_avg"], rm) < 2e-6 and rel_err(st["exp_avg_sq"], rv) < 2e-6
        assert int(st["step"]) == 5
    # state_dict layout is torch's: loads into torch.optim.AdamW and back
    topt2 = torch.optim.AdamW(theirs, lr=lr, weight_decay=wd)
    topt2.load_state_dict(opt.state_dict())
    assert rel_err(topt2.state[theirs[0]]["exp_avg"], ref[0][1]) < 2e-6


def test_adamw_many_tensors_and_grad_scale():
    """More tensors than one launch's table (48) and the optional device-side gradient scale."""
    from mae_clip_b200.train import AdamW
    g = torch.Generator().manual_seed(3)
    ps = [torch.randn(17 + i, generator=g) for i in range(120)]
    gs = [torch.randn(17 + i, generator=g) for i in range(120)]
    ours = [p.clone().cuda().requires_grad_(True) for p in ps]
    for p, gr in zip(ours, gs):
        p.grad = gr.cuda()
    AdamW(ours, lr=1e-2).step(grad_scale=torch.tensor(0.5, device="cuda"))
    for p, p0, gr in zip(ours, ps, gs):
        rp, _, _ = adamw_ref.adamw_step_ref(p0.numpy(), 0.5 * gr.numpy(), np.zeros_like(p0.numpy()), np.zeros_like(p0.numpy()),
                                            1, lr=1e-2)
        assert rel_err(p.detach(), rp) < 2e-6


def test_adamw_rejects_cpu_and_half():
    from mae_clip_b200._lib import MaeClipB200Error
    from mae_clip_b200.train import AdamW
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(MaeClipB200Error):
        AdamW([p]).step()
    q = torch.zeros(4, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    q.grad = torch.ones_like(q)
    with pytest.raises(TypeError):
        AdamW([q]).step()


class _Towers(torch.nn.Module):
    """Tiny stand-in towers (the real ones are out of scope): linear maps to the heads' input widths."""

    def __init__(self, out):
        super().__init__()
        self.lin = torch.nn.Linear(48, out)

    def forward(self, x=None, input_ids=None, attention_mask=None):
        x = x if x is not None else input_ids.float()
        return self.lin(x.reshape(x.shape[0], -1)[:, :48])


def _model(seed):
    import mae_clip_b200 as m
    torch.manual_seed(seed)
    return m.CLIPModel(image_embedding=64, text_embedding=32, image_encoder=_Towers(64), text_encoder=_Towers(32)).cuda()


def _loader(n_batches, B, seed):
    g = torch.Generator().manual_seed(seed)
    return [{"image": torch.randn(B, 48, generator=g), "input_ids": torch.randint(0, 5, (B, 48), generator=g),
             "attention_mask": torch.ones(B, 48, dtype=torch.long), "caption": ["x"] * B} for _ in range(n_batches)]


def test_train_and_valid_epoch_match_reference_loop():
    """Same batches, same init, dropout off: our loop (device-side meter, fused AdamW) against the reference's
    loop shape (main.py:51-67: loss.item() per step, torch.optim.AdamW)."""
    from mae_clip_b200.train import AdamW, get_lr, train_epoch, valid_epoch
    loader = _loader(4, 128, 1)
    a, b = _model(0), _model(0)
    a.eval(); b.eval()  # no dropout noise; gradients still flow
    oa, ob = AdamW(a.parameters(), lr=1e-3, weight_decay=1e-3), torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=1e-3)
    meter = train_epoch(a, loader, oa, None, "epoch")
    tot, cnt = 0.0, 0
    for batch in loader:  # the reference loop, restated
        batch = {k: v.cuda() for k, v in batch.items() if k != "caption"}
        loss = b(batch)
        ob.zero_grad()
        loss.backward()
        ob.step()
        tot += loss.item() * batch["image"].size(0)
        cnt += batch["image"].size(0)
    assert meter.count == cnt and abs(meter.avg - tot / cnt) < 2e-4 * abs(tot / cnt)
    for pa, pb in zip(a.parameters(), b.parameters()):
        # the two runs diverge by optimiser rounding after step 1, which re-rolls the fp16 rounding of the
        # gradient sweep's weight tiles (2.5e-4 relative, inside the 1e-3 gradient tolerance)
        assert rel_err(pa.detach(), pb.detach()) < 1e-3
    assert get_lr(oa) == 1e-3
    with torch.no_grad():
        va, = [valid_epoch(a, loader)]
    assert va.count == cnt and np.isfinite(va.avg) and "Metric" in repr(va)


def test_graphed_step_matches_eager():
    import mae_clip_b200 as m
    from mae_clip_b200.train import GraphedStep
    B = 256
    I0 = loss_ref.make_embeddings(B, 256, seed=1, scale=0.3).cuda()
    T0 = loss_ref.make_embeddings(B, 256, seed=2, scale=0.3).cuda()
    head = m.ProjectionHead(256).cuda().eval()
    params = list(head.parameters())

    def loss_fn(I, T):
        return m.clip_contrastive_loss(head(I), T, 1.0)

    gs = GraphedStep(loss_fn, [I0, T0.requires_grad_(True)], params)  # the text embeddings want a gradient too
    for seed in (5, 6):
        I = loss_ref.make_embeddings(B, 256, seed=seed, scale=0.3).cuda()
        T = loss_ref.make_embeddings(B, 256, seed=seed + 10, scale=0.3).cuda().requires_grad_(True)
        lg = gs(I, T).clone()
        graph_grads = [p.grad.clone() for p in params]
        graph_dT = gs.static_in[1].grad.clone()
        for p in params:
            p.grad = None
        le = loss_fn(I, T)
        le.backward()
        assert abs(lg.item() - le.item()) <= 1e-6 * abs(le.item())
        for gg, p in zip(graph_grads, params):
            assert rel_err(gg, p.grad) < 1e-6
        assert rel_err(graph_dT, T.grad) < 1e-6
