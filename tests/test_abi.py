"""C-ABI boundary checks that need no GPU: the library builds, loads and exports exactly the entry
points ``include/mae_clip_b200.h`` declares; the ctypes table mirrors the header; the product path
refuses CPU tensors instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mae_clip_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported(lib_built):
    declared = _declared()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", lib_built], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [n for n in declared if n not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = sorted(n for n in exported if n.startswith("mc_") and n not in declared)
    assert not extra, f"exported but not declared in the header: {extra}"


def test_ctypes_table_matches_header(lib_built):
    from mae_clip_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    h = _lib.lib()  # loads and binds every symbol
    assert h.mc_version() >= 100
    # argument counts of the ctypes table equal the header's parameter counts
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_res, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), f"{name}: header has {n} parameters, ctypes table {len(args)}"


def test_header_is_plain_c(lib_built, tmp_path):
    """The boundary is a C ABI: the header must compile as C with no CUDA/torch types."""
    c = tmp_path / "t.c"
    c.write_text('#include "mae_clip_b200.h"\nint main(void){return mc_version()==0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), "-c", str(c), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly by every public op (no silent PyTorch path)."""
    import mae_clip_b200 as m
    from mae_clip_b200._lib import MaeClipB200Error
    I = torch.randn(8, 256)
    with pytest.raises(MaeClipB200Error):
        m.clip_contrastive_loss(I, I)
    with pytest.raises(MaeClipB200Error):
        m.cross_entropy(torch.randn(4, 4), torch.rand(4, 4))
    head = m.ProjectionHead(32)
    with pytest.raises(MaeClipB200Error):
        head(torch.randn(4, 32))
    with pytest.raises(MaeClipB200Error):
        m.random_masking(torch.randn(2, 16, 8), 0.75, torch.rand(2, 16))
    with pytest.raises(MaeClipB200Error):
        m.masked_mse_loss(torch.randn(2, 4, 768), torch.randn(2, 3, 32, 32), torch.ones(2, 4))


def test_product_never_imports_oracle():
    """Nothing under mae_clip_b200/ may import, call or link the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "mae_clip_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(\.*)oracle\b", text, flags=re.M), f
                assert not re.search(r"importlib[^\n]*oracle|__import__[^\n]*oracle", text), f


def test_drop_in_surface():
    """Constructor signatures, defaults, submodule names and state_dict keys of the reference
    (CLIP.py:10-21, modules.py:55-67) are preserved."""
    import inspect

    import mae_clip_b200 as m
    head = m.ProjectionHead(2048)
    assert list(head.state_dict().keys()) == ["projection.weight", "projection.bias", "fc.weight", "fc.bias",
                                              "layer_norm.weight", "layer_norm.bias"]
    assert head.projection.weight.shape == (256, 2048) and head.dropout.p == 0.1
    assert [n for n, _ in head.named_children()] == ["projection", "gelu", "fc", "dropout", "layer_norm"]
    sig = inspect.signature(m.CLIPModel.__init__)
    assert sig.parameters["temperature"].default == 1.0
    assert sig.parameters["image_embedding"].default == 2048
    assert sig.parameters["text_embedding"].default == 768
    sig = inspect.signature(m.cross_entropy)
    assert list(sig.parameters) == ["preds", "targets", "reduction"] and sig.parameters["reduction"].default == "none"
    sig = inspect.signature(m.ProjectionHead.__init__)
    assert sig.parameters["projection_dim"].default == 256 and sig.parameters["dropout"].default == 0.1


def test_reference_state_dict_loads(golden):
    """A reference ProjectionHead checkpoint (fixture) loads into the drop-in unchanged."""
    import mae_clip_b200 as m
    z = golden("proj_head_model")
    keys = ["projection.weight", "projection.bias", "fc.weight", "fc.bias", "layer_norm.weight", "layer_norm.bias"]
    head = m.ProjectionHead(160)
    head.load_state_dict({k: torch.from_numpy(z[f"img.{k}"]) for k in keys}, strict=True)


def test_checkpoint_keys_match_reference_fixture():
    """CPU side of the checkpoint compatibility (`main.py:118-121`, `inference.py:18`): the drop-in model's state_dict
    has exactly the keys and shapes of the checkpoint the unmodified reference wrote (tests/golden/clip_checkpoint.npz);
    the scale state a ProjectionHead keeps between calls is not part of it."""
    import os
    import sys
    import numpy as np
    import mae_clip_b200 as m
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _tiny_towers as tt
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clip_checkpoint.npz"))
    ref = {k[3:]: z[k].shape for k in z.files if k.startswith("sd.")}
    model = m.CLIPModel(temperature=1.0, image_embedding=tt.IMG_DIM, text_embedding=tt.TXT_DIM,
                        image_encoder=tt.ImageTower(), text_encoder=tt.TextTower())
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert mine == {k: tuple(v) for k, v in ref.items()}
    assert not any("scale_state" in k for k in mine)
