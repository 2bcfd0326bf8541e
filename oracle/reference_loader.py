"""Import the UNMODIFIED reference denoiser for CPU baselines and parity checks -- TEST INFRASTRUCTURE, NOT PRODUCT.

The reference (rigelshysaj/LaVie) is pure Python/PyTorch and ships neither a setup.py nor its two third-party
dependencies (diffusers==0.16.0, rotary_embedding_torch), so the base contract's ``pip install --target baseline/_ref``
cannot work (no project metadata to build a wheel from; recorded in DESIGN.md).  Instead ``__graft_entry__.build()``
copies the reference's own model files, byte for byte, into the git-ignored ``baseline/_ref/`` (it still travels to the
GPU box with gpurun), and this loader puts them on ``sys.path`` together with the stand-ins for the two missing packages
(``tests/golden/shims``, written from published behaviour -- see their docstrings).  Only ``bench.py``'s CPU legs
(``cpu_baseline`` / ``--impl reference``) and ``tests/`` call this.
"""
from __future__ import annotations

import importlib
import os
import sys
from typing import Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "tests", "golden", "shims")
# reference sub-trees -> the model files the per-step denoiser needs
REF_FILES = {
    "base": ["models/__init__.py", "models/unet.py", "models/unet_blocks.py", "models/attention.py",
             "models/resnet.py"],
    "interpolation": ["models/__init__.py", "models/unet.py", "models/unet_blocks.py", "models/attention.py",
                      "models/resnet.py", "models/utils.py", "models/clip.py"],
    "vsr": ["models/__init__.py", "models/unet.py", "models/unet_blocks.py", "models/attention.py", "models/resnet.py",
            "models/temporal_module.py", "models/diffusers_attention.py", "configs/unet_3d_config.json"],
}


def populate(reference_root: str = "/root/reference") -> bool:
    """Copy the reference's model sources into baseline/_ref (build container only).  Returns False when the reference
    tree is not present (GPU box: the directory travels with the snapshot instead)."""
    import shutil
    if not os.path.isdir(reference_root):
        return False
    for tree, files in REF_FILES.items():
        for rel in files:
            src = os.path.join(reference_root, tree, rel)
            dst = os.path.join(REF_DIR, tree, rel)
            if not os.path.exists(src):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return True


def available(tree: str = "base") -> bool:
    return os.path.exists(os.path.join(REF_DIR, tree, "models", "unet.py"))


def load_reference_unet(variant: str = "base", state_dict=None):
    """Build the reference's UNet3DConditionModel (eval, CPU fp32) for ``variant`` in {"base", "interp"} -- or its
    UNet3DVSRModel for "vsr", from the reference's own vsr/configs/unet_3d_config.json -- and load ``state_dict`` with
    strict=True.  Raises FileNotFoundError when baseline/_ref is missing."""
    tree = {"base": "base", "interp": "interpolation", "vsr": "vsr"}[variant]
    if not available(tree):
        raise FileNotFoundError(f"{REF_DIR}/{tree} not found: run __graft_entry__.build() in the build container")
    for p in (SHIMS, os.path.join(REF_DIR, tree)):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        del sys.modules[name]                      # the two trees use the same package name
    mod = importlib.import_module("models.unet")
    from lavie_b200.config import BASE_CONFIG, INTERP_CONFIG
    if variant == "vsr":
        import json
        with open(os.path.join(REF_DIR, "vsr", "configs", "unet_3d_config.json")) as f:
            ref = mod.UNet3DVSRModel.from_config(json.load(f)).eval()
        if state_dict is not None:
            ref.load_state_dict(state_dict, strict=True)
        return ref
    if variant == "base":
        cfg = BASE_CONFIG.to_dict()
    else:
        cfg = INTERP_CONFIG.to_dict()
        cfg["use_first_frame"] = True              # what from_pretrained_2d sets for copy_no_mask, unet.py:487-507
    ref = mod.UNet3DConditionModel.from_config(cfg).eval()
    if state_dict is not None:
        ref.load_state_dict(state_dict, strict=True)
    return ref
