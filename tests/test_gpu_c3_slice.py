"""BASELINE config 3 as a parity case (not a bench line): the hot-path slice of a ViT-B/16 MAE + DistilBERT step at
batch 256 with bf16 activations - patch tokens (256, 196, 768) -> random masking at ratio 0.75 -> mean-pooled kept
tokens -> image ProjectionHead(768); text CLS features (256, 768) -> text ProjectionHead(768); contrastive soft-target
loss; decoder prediction (256, 196, 768) bf16 + images -> normalised-pixel masked MSE; one backward through all of it.
The towers themselves (transformer blocks) are out of scope and are not run.  Oracle: the same
composition from oracle/ on CPU in fp32 from the same bf16-rounded inputs.  Tolerances: mask indices bit-exact; the
kernels compute in fp32 from bf16 inputs, so the fp32 bars (loss 1e-4, gradients 1e-3) apply to fp32 outputs and
gradients that are themselves stored in bf16 are compared after the same rounding (bf16 bar: 8e-3)."""
import pytest
import torch

from conftest import rel_err
from oracle import loss_ref, mae_ref, proj_head_ref

pytestmark = pytest.mark.gpu


def test_c3_hot_path_slice_bf16():
    import mae_clip_b200 as m
    B, L, Dm, ratio, tau = 256, 196, 768, 0.75, 1.0
    g = torch.Generator().manual_seed(33)
    tokens = torch.randn(B, L, Dm, generator=g).bfloat16()
    noise = torch.rand(B, L, generator=g)
    text_feat = torch.randn(B, 768, generator=g).bfloat16()
    pred = (torch.randn(B, L, 768, generator=g) * 0.5).bfloat16()
    imgs = torch.randn(B, 3, 224, 224, generator=g)
    torch.manual_seed(7)
    head_i, head_t = m.ProjectionHead(768).cuda().eval(), m.ProjectionHead(768).cuda().eval()

    # ---- B200 path
    tok_c = tokens.cuda().requires_grad_(True)
    txt_c = text_feat.cuda().requires_grad_(True)
    pred_c = pred.cuda().requires_grad_(True)
    x_masked, mask, ids_restore = m.random_masking(tok_c, ratio, noise.cuda())
    img_feat = x_masked.float().mean(dim=1)                      # stand-in for the encoder + pooling (stock torch)
    e_i, e_t = head_i(img_feat) * 0.1, head_t(txt_c.float()) * 0.1
    l_clip = m.clip_contrastive_loss(e_i, e_t, tau)
    l_mae = m.masked_mse_loss(pred_c, imgs.cuda(), mask, patch_size=16, norm_pix_loss=True)
    (l_clip + l_mae).backward()

    # ---- oracle, fp32 on CPU from the same bf16-rounded inputs
    tok_r = tokens.float().requires_grad_(True)
    txt_r = text_feat.float().requires_grad_(True)
    pred_r = pred.float().requires_grad_(True)
    xm_r, mask_r, restore_r, _ = mae_ref.random_masking_ref(tok_r, ratio, noise)
    pi = [p.detach().cpu() for p in head_i.parameters()]
    pt = [p.detach().cpu() for p in head_t.parameters()]
    ei_r = proj_head_ref.proj_head_ref(xm_r.mean(dim=1), *pi) * 0.1
    et_r = proj_head_ref.proj_head_ref(txt_r, *pt) * 0.1
    lc_r = loss_ref.clip_loss_ref(ei_r, et_r, tau)
    lm_r = mae_ref.masked_mse_ref(pred_r, imgs, mask_r)
    (lc_r + lm_r).backward()

    assert torch.equal(ids_restore.cpu(), restore_r) and torch.equal(mask.cpu(), mask_r)      # indices: bit-exact
    assert torch.equal(x_masked.detach().cpu().float(), xm_r.detach())                         # gather: bit-exact
    assert abs(l_clip.item() - lc_r.item()) < 1e-4 * abs(lc_r.item())
    assert abs(l_mae.item() - lm_r.item()) < 1e-4 * abs(lm_r.item())
    # gradients stored in bf16 (token / text / pred gradients take the dtype of their inputs)
    assert tok_c.grad.dtype == torch.bfloat16 and pred_c.grad.dtype == torch.bfloat16
    assert rel_err(tok_c.grad.float(), tok_r.grad) < 8e-3
    assert rel_err(txt_c.grad.float(), txt_r.grad) < 8e-3
    assert rel_err(pred_c.grad.float(), pred_r.grad) < 8e-3
    # positions that were masked out of the encoder input get exactly zero gradient
    assert (tok_c.grad.float() * mask.unsqueeze(-1)).abs().max().item() == 0.0
