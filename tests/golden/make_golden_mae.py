"""Golden fixtures for the MAE rows (M1-M3) from a PUBLISHED implementation.

``/root/reference`` holds no MAE code, so the MAE oracle cannot be pinned on the
reference itself.  The formulation BASELINE.json names (noise argsort, keep / restore gather, normalised-pixel
masked-patch MSE) is the one of "Masked Autoencoders Are Scalable Vision Learners"; its implementation in
``transformers`` (``ViTMAEEmbeddings.random_masking``, ``ViTMAEForPreTraining.patchify / forward_loss``;
version recorded in the fixture) is importable here, so its outputs pin ``oracle/mae_ref.py`` and the CUDA path:

    python tests/golden/make_golden_mae.py        # writes tests/golden/mae_hf.npz

Every ``ref_`` array is an output of the transformers code run unmodified on CPU; the other keys are seeded inputs.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import transformers
from transformers import ViTMAEConfig, ViTMAEForPreTraining

HERE = os.path.dirname(os.path.abspath(__file__))


def tiny_model(image_size, mask_ratio, norm_pix):
    cfg = ViTMAEConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64,
                       decoder_hidden_size=32, decoder_num_hidden_layers=1, decoder_num_attention_heads=2,
                       decoder_intermediate_size=64, image_size=image_size, patch_size=16, num_channels=3,
                       mask_ratio=mask_ratio, norm_pix_loss=norm_pix)
    return ViTMAEForPreTraining(cfg).eval()


def main():
    out = {"transformers_version": np.array(transformers.__version__)}
    g = torch.Generator().manual_seed(2024)
    # ---- M1: random masking at the reference patch count (196) and a ragged one, ratios of config C5
    for tag, (N, L, D) in {"l196": (6, 196, 24), "l50": (4, 50, 8)}.items():
        x = torch.randn(N, L, D, generator=g)
        noise = torch.rand(N, L, generator=g)
        out[f"mask.{tag}.x"], out[f"mask.{tag}.noise"] = x.numpy(), noise.numpy()
        for ratio in (0.5, 0.6, 0.75, 0.9):
            m = tiny_model(224, ratio, True)
            with torch.no_grad():
                xm, mask, ids_restore = m.vit.embeddings.random_masking(x, noise)
            k = f"mask.{tag}.r{int(ratio * 100)}"
            out[k + ".ref_x_masked"], out[k + ".ref_mask"], out[k + ".ref_ids_restore"] = \
                xm.numpy(), mask.numpy(), ids_restore.numpy()
    # ---- M2 + M3: patchify, normalised-pixel target (through the loss), masked MSE and its gradient
    for tag, size in {"s64": 64, "s224": 224}.items():
        N = 3 if size == 64 else 1
        L = (size // 16) ** 2
        imgs = torch.randn(N, 3, size, size, generator=g)
        pred = torch.randn(N, L, 768, generator=g) * 0.7
        noise = torch.rand(N, L, generator=g)
        out[f"mse.{tag}.imgs"], out[f"mse.{tag}.pred"] = imgs.numpy(), pred.numpy()
        for norm_pix in ((True, False) if size == 64 else (True,)):  # keep the fixture small: 224 px once
            m = tiny_model(size, 0.75, norm_pix)
            with torch.no_grad():
                _, mask, _ = m.vit.embeddings.random_masking(torch.zeros(N, L, 4), noise)
                patches = m.patchify(imgs)
            p = pred.clone().requires_grad_(True)
            loss = m.forward_loss(imgs, p, mask)
            loss.backward()
            k = f"mse.{tag}.np{int(norm_pix)}"
            out[k + ".mask"] = mask.numpy()
            out[k + ".ref_loss"] = loss.detach().numpy()
            out[k + ".ref_dpred"] = p.grad.numpy()
            if norm_pix and size == 64:
                out[f"mse.{tag}.ref_patchify"] = patches.numpy()
    np.savez_compressed(os.path.join(HERE, "mae_hf.npz"), **out)
    print("wrote mae_hf.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
