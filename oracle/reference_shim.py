"""Import the UNMODIFIED reference (``/root/reference``) for golden generation.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Only usable in the dev
container (``/root/reference`` does not exist on the GPU box); nothing run with
``-m gpu``, ``smoke()`` or ``bench.py`` may call this.

Recipe (SURVEY.md section 8 c): ``modules.py:3`` imports ``timm`` which is absent
here, so a stub module is registered AFTER importing ``transformers`` (whose
availability probe chokes on a spec-less stub).  The stub's ``create_model``
is never on the hot path: fixtures only use ``ProjectionHead`` and
``cross_entropy`` plus a CLIPModel with identity encoders.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

REFERENCE_DIR = os.environ.get("MAE_CLIP_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "CLIP.py"))


def load():
    """Returns (CLIP module, modules module) of the reference, imported as-is."""
    if not available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_DIR}")
    import torch  # noqa: F401
    from transformers import DistilBertModel  # noqa: F401  (must precede the timm stub)

    if "timm" not in sys.modules:
        stub = types.ModuleType("timm")
        stub.__spec__ = importlib.machinery.ModuleSpec("timm", None)

        def create_model(*_a, **_k):
            import torch.nn as nn
            return nn.Identity()

        stub.create_model = create_model
        sys.modules["timm"] = stub
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    saved = {k: sys.modules.pop(k) for k in ("config", "modules", "CLIP") if k in sys.modules}
    try:
        import config as ref_cfg  # noqa: F401
        import modules as ref_modules
        import CLIP as ref_clip
    finally:
        for k in ("config", "modules", "CLIP"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
        if REFERENCE_DIR in sys.path:
            sys.path.remove(REFERENCE_DIR)
    return ref_clip, ref_modules
