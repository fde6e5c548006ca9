"""Whole-step timings around the hot path (SURVEY.md section 8 d, configs C1 and C3): how much of a real step the
accelerated path is.  The towers are STOCK PyTorch (out of scope): torchvision blocks and a
random-init DistilBERT; only the heads, the contrastive loss, the MAE masking / un-shuffle / masked MSE and the
optimiser step run in libmae_clip_b200.so.

  C1  CLIPModel (ResNet-50 image tower + DistilBERT, 256-d heads), batch 32, 224 px, fp32: one fwd + bwd + AdamW
      step on the B200, and the same step of the reference composition on this box's CPU cores (oracle port).
  C3  ViT-B/16 MAE encoder at mask ratio 0.75 + 8-block decoder + DistilBERT text tower, batch 256, bf16 autocast
      for the towers: one fwd + bwd step; the hot-path share is measured by timing the same step with the
      hot-path calls replaced by no-ops of the same shapes.

    python tools/full_step_bench.py [--json out.json] [--skip-cpu]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch import nn  # noqa: E402

import mae_clip_b200 as m  # noqa: E402
from mae_clip_b200.train import AdamW  # noqa: E402

dev = torch.device("cuda")
out = {}


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


# ------------------------------------------------------------------------------------------------ C1
def c1():
    torch.manual_seed(0)
    B = 32
    model = m.CLIPModel(image_encoder=m.ImageEncoder(pretrained=False), text_encoder=m.TextEncoder(pretrained=False)).to(dev)
    model.train()
    opt = AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-3)
    batch = {"image": torch.randn(B, 3, 224, 224, device=dev), "input_ids": torch.randint(5, 300, (B, 25), device=dev),
             "attention_mask": torch.ones(B, 25, dtype=torch.long, device=dev)}

    def step():
        loss = model(batch)
        opt.zero_grad()
        loss.backward()
        opt.step()
    ms = timeit(step)
    # hot path alone on the same shapes: heads + loss fwd/bwd + their optimiser step
    fi = torch.randn(B, 2048, device=dev, requires_grad=True)
    ft = torch.randn(B, 768, device=dev)
    hp = list(model.image_projection.parameters()) + list(model.text_projection.parameters())
    hopt = AdamW(hp, lr=1e-3, weight_decay=1e-3)

    def hot():
        loss = m.clip_contrastive_loss(model.image_projection(fi), model.text_projection(ft), 1.0)
        hopt.zero_grad()
        loss.backward()
        hopt.step()
    hot_ms = timeit(hot)
    out["c1_full_model_step_B32_fp32"] = {"b200_ms": ms, "samples_per_s": B / (ms * 1e-3), "hot_path_ms": hot_ms,
                                          "hot_path_share": hot_ms / ms}
    print("C1 B200:", out["c1_full_model_step_B32_fp32"], flush=True)
    return model


def c1_cpu():
    """The reference composition on the host cores in plain torch ops (same towers; heads and loss written out as
    modules.py:69-76 and CLIP.py:34-43 do; torch AdamW).  A timing baseline only - the checker lives in oracle/."""
    import torch.nn.functional as F
    import torchvision
    from transformers import DistilBertConfig, DistilBertModel

    def head(x, h):
        p = F.linear(x, h.projection.weight, h.projection.bias)
        y = F.dropout(F.linear(F.gelu(p), h.fc.weight, h.fc.bias), 0.1, True)
        return F.layer_norm(y + p, (256,), h.layer_norm.weight, h.layer_norm.bias, 1e-5)

    def clip_loss(ie, te, tau):
        logits = (te @ ie.T) / tau
        targets = F.softmax((ie @ ie.T + te @ te.T) / 2 * tau, dim=-1)
        tl = (-targets * F.log_softmax(logits, dim=-1)).sum(1)
        il = (-targets.T * F.log_softmax(logits.T, dim=-1)).sum(1)
        return ((il + tl) / 2.0).mean()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    B = 32
    img = torchvision.models.resnet50(weights=None)
    img.fc = nn.Identity()
    txt = DistilBertModel(DistilBertConfig())
    for p in txt.parameters():
        p.requires_grad = False                                 # modules.py:35: the text tower is frozen
    hi, ht = m.ProjectionHead(2048), m.ProjectionHead(768)      # parameter containers only; arithmetic below is the oracle's
    params = [p for p in list(img.parameters()) + list(hi.parameters()) + list(ht.parameters())]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-3)
    batch = {"image": torch.randn(B, 3, 224, 224), "input_ids": torch.randint(5, 300, (B, 25)),
             "attention_mask": torch.ones(B, 25, dtype=torch.long)}
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        fi = img(batch["image"])
        ft = txt(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).last_hidden_state[:, 0, :]
        loss = clip_loss(head(fi, hi), head(ft, ht), 1.0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        ts.append(time.perf_counter() - t0)
    ms = min(ts[1:]) * 1e3
    out["c1_full_model_step_B32_fp32"]["cpu_port_ms"] = ms
    out["c1_full_model_step_B32_fp32"]["cpu_cores"] = os.cpu_count()
    print("C1 CPU port:", ms, "ms on", os.cpu_count(), "cores", flush=True)


# ------------------------------------------------------------------------------------------------ C3
class Blocks(nn.Module):
    def __init__(self, n, dim, heads):
        super().__init__()
        from torchvision.models.vision_transformer import EncoderBlock
        self.layers = nn.ModuleList([EncoderBlock(heads, dim, 4 * dim, 0.0, 0.0) for _ in range(n)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)

    def forward(self, x):
        for blk in self.layers:
            x = blk(x)
        return self.norm(x)


class MaeClipC3(nn.Module):
    """ViT-B/16 MAE (12 x 768 encoder on the kept tokens, 8 x 512 decoder on all tokens) + DistilBERT CLS."""

    def __init__(self, hot=True):
        super().__init__()
        from transformers import DistilBertConfig, DistilBertModel
        self.hot = hot
        self.patch = nn.Conv2d(3, 768, 16, 16)
        self.pos = nn.Parameter(torch.zeros(1, 196, 768))
        self.enc = Blocks(12, 768, 12)
        self.dec_embed = nn.Linear(768, 512)
        self.mask_token = nn.Parameter(torch.zeros(512))
        self.dec_pos = nn.Parameter(torch.zeros(1, 196, 512))
        self.dec = Blocks(8, 512, 16)
        self.dec_pred = nn.Linear(512, 768)
        self.text = DistilBertModel(DistilBertConfig())
        for p in self.text.parameters():
            p.requires_grad = False
        self.image_projection = m.ProjectionHead(768)
        self.text_projection = m.ProjectionHead(768)

    def forward(self, batch, noise):
        imgs = batch["image"]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            tok = self.patch(imgs).flatten(2).transpose(1, 2) + self.pos          # (B, 196, 768) bf16
            if self.hot:
                x, mask, ids_restore = m.random_masking(tok, 0.75, noise)
            else:  # same shapes, no hot-path work: first 49 tokens, constant mask
                x, mask = tok[:, :49], torch.ones(tok.shape[0], 196, device=tok.device)
                ids_restore = None
            x = self.enc(x)
            feat_i = x.float().mean(dim=1)
            d = self.dec_embed(x)
            if self.hot:
                d = m.restore_tokens(d, self.mask_token, ids_restore)
            else:
                d = torch.cat([d, d, d, d], dim=1)                                  # (B, 196, 512) stand-in
            pred = self.dec_pred(self.dec(d + self.dec_pos))                       # (B, 196, 768) bf16
            feat_t = self.text(input_ids=batch["input_ids"],
                               attention_mask=batch["attention_mask"]).last_hidden_state[:, 0, :].float()
        if self.hot:
            e_i, e_t = self.image_projection(feat_i), self.text_projection(feat_t)
            return m.clip_contrastive_loss(e_i, e_t, 1.0) + m.masked_mse_loss(pred, imgs, mask)
        # no-op stand-ins that keep every tower's backward alive
        return feat_i.mean() + feat_t.mean() + pred.float().mean()


def c3():
    torch.manual_seed(0)
    B = 256
    batch = {"image": torch.randn(B, 3, 224, 224, device=dev), "input_ids": torch.randint(5, 300, (B, 25), device=dev),
             "attention_mask": torch.ones(B, 25, dtype=torch.long, device=dev)}
    noise = torch.rand(B, 196, device=dev)
    res = {}
    for hot in (True, False):
        model = MaeClipC3(hot=hot).to(dev).train()

        def step():
            for p in model.parameters():
                p.grad = None
            model(batch, noise).backward()
        res[hot] = timeit(step, iters=5, warm=2)
        del model
        torch.cuda.empty_cache()
    out["c3_vitb16_mae_distilbert_step_B256_bf16"] = {
        "b200_ms_with_hot_path": res[True], "b200_ms_towers_only": res[False],
        "hot_path_ms": res[True] - res[False], "hot_path_share": (res[True] - res[False]) / res[True],
        "samples_per_s": B / (res[True] * 1e-3),
        "note": "towers are stock torch (torchvision EncoderBlock, transformers DistilBERT) under bf16 autocast; "
                "hot path = random masking, un-shuffle, two heads, contrastive loss, masked MSE (fwd + bwd)"}
    print("C3:", out["c3_vitb16_mae_distilbert_step_B256_bf16"], flush=True)


if __name__ == "__main__":
    c1()
    if "--skip-cpu" not in sys.argv:
        c1_cpu()
    c3()
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
