"""Checkpoint fixture from the UNMODIFIED reference (dev container only: needs /root/reference).

    python tests/golden/make_golden_ckpt.py

Builds the reference's ``CLIPModel`` (``CLIP.py:10-21``) with its own ``ImageEncoder`` / ``TextEncoder`` /
``ProjectionHead`` classes; only the two THIRD-PARTY constructors they call are pointed at tiny stand-ins
(``tests/_tiny_towers.py``) so that the checkpoint is committable: ``timm.create_model`` (timm is absent here anyway)
and ``DistilBertConfig()``.  Saves what ``main.py:118-121`` saves (``model.state_dict()``), a batch, and the loss the
reference computes on it in eval mode (``CLIP.py:23-43``), as ``tests/golden/clip_checkpoint.npz``."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _tiny_towers as tt  # noqa: E402
from oracle import reference_shim  # noqa: E402


def main():
    torch.manual_seed(0)
    ref_clip, ref_modules = reference_shim.load()
    sys.modules["timm"].create_model = lambda *a, **k: tt.tiny_image_model()
    ref_modules.timm.create_model = sys.modules["timm"].create_model
    ref_modules.DistilBertConfig = tt.tiny_distilbert_config
    # CLIP.py:17-18 build the towers with their defaults, which need the network (pretrained=True): random init instead
    d = list(ref_modules.TextEncoder.__init__.__defaults__)
    ref_modules.TextEncoder.__init__.__defaults__ = (d[0], False, d[2])
    di = list(ref_modules.ImageEncoder.__init__.__defaults__)
    ref_modules.ImageEncoder.__init__.__defaults__ = (di[0], False, di[2])
    model = ref_clip.CLIPModel(temperature=1.0, image_embedding=tt.IMG_DIM, text_embedding=tt.TXT_DIM)
    model.eval()
    g = torch.Generator().manual_seed(1)
    batch = {"image": torch.randn(6, 3, 32, 32, generator=g), "input_ids": torch.randint(5, 300, (6, 10), generator=g),
             "attention_mask": torch.ones(6, 10, dtype=torch.long)}
    batch["attention_mask"][2, 7:] = 0
    with torch.no_grad():
        loss = model(batch)
        img_emb = model.image_projection(model.image_encoder(batch["image"]))
    out = {"sd." + k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    out.update({"batch.image": batch["image"].numpy(), "batch.input_ids": batch["input_ids"].numpy(),
                "batch.attention_mask": batch["attention_mask"].numpy(), "ref_eval_loss": np.float64(loss.item()),
                "ref_image_embeddings": img_emb.numpy()})
    path = os.path.join(HERE, "clip_checkpoint.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(model.state_dict())} state_dict entries, eval loss {loss.item():.6f}, "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
