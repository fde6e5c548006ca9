"""GPU parity of the MAE branch (rows M1-M3 of SURVEY.md section 8) against ``oracle/mae_ref.py``.

The reference tree has no MAE code, so the oracle is the defining restatement ("parity
unpinned"); indices, masks and gathered rows must match it BIT-EXACTLY, the masked MSE within
fp32 tolerance."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import mae_ref

pytestmark = pytest.mark.gpu


def _noise(N, L, seed, ties=True):
    g = torch.Generator().manual_seed(seed)
    noise = torch.rand(N, L, generator=g)
    if ties and N > 1 and L > 12:
        noise[0, 5] = noise[0, L - 3]          # a tie far apart
        noise[1, :10] = 0.25                   # a run of ties
        noise[N - 1, :] = 0.5                  # a fully tied row -> identity permutation
    return noise


@pytest.mark.parametrize("N,L,D", [(64, 196, 768), (3, 196, 768), (1, 1, 8), (5, 7, 12), (2, 1024, 16), (130, 196, 64)])
@pytest.mark.parametrize("ratio", [0.5, 0.6, 0.75, 0.9])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_random_masking_bit_exact(N, L, D, ratio, dtype):
    import mae_clip_b200 as m
    g = torch.Generator().manual_seed(N * L)
    x = torch.randn(N, L, D, generator=g).to(dtype)
    noise = _noise(N, L, seed=N + L)
    xm_ref, mask_ref, restore_ref, keep_ref = mae_ref.random_masking_ref(x, ratio, noise)
    xm, mask, restore, keep = m.random_masking_with_ids(x.cuda(), ratio, noise.cuda())
    assert restore.dtype == torch.int64 and mask.dtype == torch.float32 and xm.dtype == dtype
    assert torch.equal(restore.cpu(), restore_ref)
    assert torch.equal(keep.cpu(), keep_ref)
    assert torch.equal(mask.cpu(), mask_ref)
    assert torch.equal(xm.cpu(), xm_ref)
    # independent numpy restatement of the index contract
    k2, r2, m2 = mae_ref.random_masking_numpy(noise.numpy(), ratio)
    assert np.array_equal(restore.cpu().numpy(), r2) and np.array_equal(mask.cpu().numpy(), m2)


def test_random_masking_full_size_properties():
    """C5 upper end (N=1024, L=196, D=768): permutation validity and round trip, no CPU oracle."""
    import mae_clip_b200 as m
    N, L, D = 1024, 196, 768
    x = torch.randn(N, L, D, device="cuda")
    noise = torch.rand(N, L, device="cuda")
    for ratio in (0.5, 0.75, 0.9):
        keep_n = mae_ref.len_keep_for(L, ratio)
        xm, mask, restore, keep = m.random_masking_with_ids(x, ratio, noise)
        assert xm.shape == (N, keep_n, D)
        assert torch.equal(torch.sort(restore, dim=1).values, torch.arange(L, device="cuda").expand(N, L))
        assert mask.sum().item() == N * (L - keep_n)
        # kept noise values are the len_keep smallest, in ascending order
        kept_noise = torch.gather(noise, 1, keep)
        assert (kept_noise[:, 1:] >= kept_noise[:, :-1]).all()
        assert (kept_noise.max(1).values <= torch.where(mask.bool(), noise, torch.full_like(noise, 2.0)).min(1).values).all()
        assert torch.equal(xm, torch.gather(x, 1, keep.unsqueeze(-1).expand(-1, -1, D)))
        full = m.restore_tokens(xm, torch.zeros(D, device="cuda"), restore)
        assert torch.equal(full, x * (1 - mask).unsqueeze(-1))


def test_random_masking_backward():
    import mae_clip_b200 as m
    N, L, D = 8, 196, 64
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, L, D, generator=g)
    noise = torch.rand(N, L, generator=g)
    go = torch.randn(N, 49, D, generator=g)
    xr = x.clone().requires_grad_(True)
    mae_ref.random_masking_ref(xr, 0.75, noise)[0].backward(go)
    xc = x.cuda().requires_grad_(True)
    m.random_masking(xc, 0.75, noise.cuda())[0].backward(go.cuda())
    assert torch.equal(xc.grad.cpu(), xr.grad)


@pytest.mark.parametrize("N,HW,p", [(4, 224, 16), (2, 32, 16), (3, 48, 8), (1, 16, 16)])
@pytest.mark.parametrize("norm_pix", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_masked_mse_vs_oracle(N, HW, p, norm_pix, dtype):
    import mae_clip_b200 as m
    g = torch.Generator().manual_seed(N + HW)
    L = (HW // p) ** 2
    imgs = torch.randn(N, 3, HW, HW, generator=g)
    pred = torch.randn(N, L, p * p * 3, generator=g).to(dtype)
    noise = torch.rand(N, L, generator=g)
    mask = mae_ref.random_masking_numpy(noise.numpy(), 0.75)[2]
    mask = torch.from_numpy(mask)
    if mask.sum() == 0:
        mask[:, 0] = 1
    loss_ref_, dpred_ref = mae_ref.masked_mse_fwd_bwd_ref(pred.float(), imgs, mask, p, norm_pix)
    pc = pred.cuda().requires_grad_(True)
    loss = m.masked_mse_loss(pc, imgs.cuda(), mask.cuda(), p, norm_pix)
    (loss * 2.0).backward()
    assert abs(loss.item() - loss_ref_.item()) < 1e-5 * abs(loss_ref_.item())
    tol = 1e-5 if dtype == torch.float32 else 8e-3  # bf16 gradient storage
    assert rel_err(pc.grad.float(), 2.0 * dpred_ref) < tol
    assert pc.grad.dtype == dtype
    # unmasked patches get exactly zero gradient
    assert (pc.grad.float().cpu()[mask == 0] == 0).all()


def test_masked_mse_full_size_properties():
    """C5: N=256, 224px, p=16: loss of pred == target is 0; loss is linear in the mask weights."""
    import mae_clip_b200 as m
    N = 256
    imgs = torch.randn(N, 3, 224, 224, device="cuda")
    target = m.patchify(imgs, 16, norm_pix=True)
    ref_t = mae_ref.norm_pix_target_ref(imgs[:2].cpu(), 16)
    assert rel_err(target[:2], ref_t) < 1e-5
    assert torch.equal(m.patchify(imgs[:2], 16, norm_pix=False).cpu(), mae_ref.patchify_ref(imgs[:2].cpu(), 16))
    mask = (torch.rand(N, 196, device="cuda") < 0.75).float()
    assert m.masked_mse_loss(target, imgs, mask).item() < 1e-10
    pred = target + 0.5
    assert abs(m.masked_mse_loss(pred, imgs, mask).item() - 0.25) < 1e-5
    pred = torch.randn_like(target)
    la = m.masked_mse_loss(pred, imgs, mask)
    per_patch = ((pred - target) ** 2).mean(-1)
    assert abs(la.item() - ((per_patch * mask).sum() / mask.sum()).item()) < 1e-5 * la.item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,L,D,ratio", [(8, 196, 768, 0.75), (3, 50, 64, 0.5), (5, 196, 512, 0.9), (2, 16, 8, 0.0)])
def test_restore_tokens_forward_backward(dtype, N, L, D, ratio):
    """Decoder-side un-shuffle (SURVEY 8 f rank 2) against oracle/mae_ref.restore_tokens_ref (cat mask tokens +
    gather through ids_restore, as in the published MAE decoder): forward bit-exact, gradients of the kept tokens
    bit-exact (a copy) and of the mask token to fp32-accumulation accuracy."""
    import mae_clip_b200 as m
    g = torch.Generator().manual_seed(N * 100 + L)
    noise = torch.rand(N, L, generator=g)
    keep = mae_ref.len_keep_for(L, ratio)
    ids_restore = torch.argsort(torch.argsort(noise, dim=1, stable=True), dim=1, stable=True)
    xk = torch.randn(N, keep, D, generator=g).to(dtype)
    tok = torch.randn(D, generator=g).to(dtype)
    go = torch.randn(N, L, D, generator=g).to(dtype)

    def ref(xk, tok):
        xk = xk.clone().requires_grad_(True)
        tok = tok.clone().requires_grad_(True)
        out = mae_ref.restore_tokens_ref(xk, tok, ids_restore)
        out.backward(go.to(out.dtype))
        return out.detach(), xk.grad, tok.grad

    o_ref, dx_ref, dt_ref = ref(xk.float(), tok.float())
    xc, tc = xk.cuda().requires_grad_(True), tok.cuda().requires_grad_(True)
    out = m.restore_tokens(xc, tc, ids_restore.cuda())
    out.backward(go.cuda())
    assert out.dtype == dtype and torch.equal(out.detach().cpu().float(), o_ref)
    assert torch.equal(xc.grad.cpu().float(), dx_ref)
    tol = 1e-6 if dtype == torch.float32 else 8e-3  # bf16: the result is rounded once to bf16
    assert rel_err(tc.grad.float(), dt_ref) < tol


# ------------------------------------------------------------------ golden vectors of a published implementation
def test_mae_cuda_matches_published_vitmae_fixture(golden):
    """CUDA path against tests/golden/mae_hf.npz (outputs of transformers' ViTMAE run unmodified, see
    tests/golden/make_golden_mae.py): masking indices, mask and gathered tokens bit-exact; patchify bit-exact;
    normalised-pixel masked MSE and its gradient within fp32 accumulation accuracy."""
    import mae_clip_b200 as m
    z = golden("mae_hf")
    for tag in ("l196", "l50"):
        x, noise = torch.from_numpy(z[f"mask.{tag}.x"]).cuda(), torch.from_numpy(z[f"mask.{tag}.noise"]).cuda()
        for r in (0.5, 0.6, 0.75, 0.9):
            k = f"mask.{tag}.r{int(r * 100)}"
            xm, mask, restore = m.random_masking(x, r, noise)
            assert np.array_equal(restore.cpu().numpy(), z[k + ".ref_ids_restore"])
            assert np.array_equal(mask.cpu().numpy(), z[k + ".ref_mask"])
            assert np.array_equal(xm.cpu().numpy(), z[k + ".ref_x_masked"])
    imgs64 = torch.from_numpy(z["mse.s64.imgs"]).cuda()
    assert np.array_equal(m.patchify(imgs64).cpu().numpy(), z["mse.s64.ref_patchify"])
    for tag, nps in (("s64", (1, 0)), ("s224", (1,))):
        imgs = torch.from_numpy(z[f"mse.{tag}.imgs"]).cuda()
        for npx in nps:
            k = f"mse.{tag}.np{npx}"
            pred = torch.from_numpy(z[f"mse.{tag}.pred"]).cuda().requires_grad_(True)
            mask = torch.from_numpy(z[k + ".mask"]).cuda()
            loss = m.masked_mse_loss(pred, imgs, mask, patch_size=16, norm_pix_loss=bool(npx))
            loss.backward()
            ref = float(z[k + ".ref_loss"])
            assert abs(loss.item() - ref) < 1e-5 * abs(ref)
            assert rel_err(pred.grad, z[k + ".ref_dpred"]) < 1e-5


def test_random_masking_ratio_one_backward_is_zero():
    """mask_ratio = 1.0 keeps nothing: forward returns empty tokens, backward a zero gradient (it used to hand a NULL
    pointer of the empty gradient to the C ABI)."""
    import mae_clip_b200 as m
    x = torch.randn(3, 16, 32, device="cuda", requires_grad=True)
    noise = torch.rand(3, 16, device="cuda")
    xm, mask, restore = m.random_masking(x, 1.0, noise)
    assert xm.shape == (3, 0, 32) and bool((mask == 1).all())
    (xm.sum() + 0.0 * x.sum()).backward()
    assert torch.equal(x.grad, torch.zeros_like(x))
