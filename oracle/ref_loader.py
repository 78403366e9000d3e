"""TEST INFRASTRUCTURE — import the reference's own hot-path modules by file path.

Only usable where ``/root/reference`` exists (the build container); the GPU box never
has it, so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.  It is
used by ``tests/golden/make_golden.py`` to produce the committed golden vectors and by
CPU tests (skipped when the reference is absent) that pin ``oracle.chada_oracle`` to
the live reference.

``import src`` fails in this image (Lightning / timm / omegaconf / hydra are absent), so
the three self-contained files are loaded with stub parent packages and ``DINOHead`` /
``trunc_normal_`` are extracted from their modules' ASTs (SURVEY.md §8c recipe).  No
reference source is copied into this repository.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("CHADAVIT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src/backbones/vit/chada_vit.py"))


def _extract(path: str, names: set, glb: dict) -> None:
    with open(path) as f:
        tree = ast.parse(f.read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    mod = ast.Module(body=body, type_ignores=[])
    exec(compile(mod, path, "exec"), glb)


_cache: dict = {}


def load() -> types.SimpleNamespace:
    """Returns a namespace with ChAdaViT, chada_vit, DINOHead, DINOLoss, MomentumUpdater,
    initialize_momentum_params — the reference's own classes."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import math
    from typing import Any, List

    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    # stub packages so `from src.utils.misc import trunc_normal_` resolves
    saved = {k: sys.modules.get(k) for k in ("src", "src.utils", "src.utils.misc")}
    pkg = types.ModuleType("src"); pkg.__path__ = []
    utils = types.ModuleType("src.utils"); utils.__path__ = []
    misc = types.ModuleType("src.utils.misc")
    misc.__dict__.update({"math": math, "torch": torch})
    _extract(os.path.join(REF_ROOT, "src/utils/misc.py"), {"_no_grad_trunc_normal_", "trunc_normal_"}, misc.__dict__)
    sys.modules.update({"src": pkg, "src.utils": utils, "src.utils.misc": misc})
    try:
        def by_path(name, rel):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            return m

        chada = by_path("_ref_chada_vit", "src/backbones/vit/chada_vit.py")
        loss = by_path("_ref_dino_loss", "src/losses/dino.py")
        mom = by_path("_ref_momentum", "src/utils/momentum.py")
        glb = {"torch": torch, "nn": nn, "F": F, "Any": Any, "List": List,
               "trunc_normal_": misc.trunc_normal_}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _extract(os.path.join(REF_ROOT, "src/methods/dino.py"), {"DINOHead"}, glb)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ns = types.SimpleNamespace(
        ChAdaViT=chada.ChAdaViT, chada_vit=chada.chada_vit,
        TransformerEncoderLayer=chada.TransformerEncoderLayer, TokenLearner=chada.TokenLearner,
        DINOHead=glb["DINOHead"], DINOLoss=loss.DINOLoss,
        MomentumUpdater=mom.MomentumUpdater, initialize_momentum_params=mom.initialize_momentum_params,
    )
    _cache["ns"] = ns
    return ns
