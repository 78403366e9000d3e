"""Loading reference checkpoints into the B200 modules.

The reference saves Lightning checkpoints whose ``state_dict`` holds the whole DINO module (``backbone.*``,
``momentum_backbone.*``, ``head.*`` ...; src/utils/checkpointer.py:132-147) and the evaluation drivers load only the backbone
after rewriting keys (main_linear.py:103-110, main_knn.py:196).  Parameter names and shapes of chadavit_b200's modules are the
reference's, so the tensors load as they are."""
from __future__ import annotations

from typing import Dict, Union

import torch
from torch import nn


def rewrite_backbone_keys(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """main_linear.py:103-110: 'encoder' -> 'backbone', strip 'backbone.', drop every original key."""
    state = dict(state)
    for k in list(state.keys()):
        if "encoder" in k:
            state[k.replace("encoder", "backbone")] = state[k]
        if "backbone" in k:
            state[k.replace("backbone.", "")] = state[k]
        del state[k]
    return state


def load_pretrained_backbone(backbone: nn.Module, ckpt: Union[str, Dict]) -> "torch.nn.modules.module._IncompatibleKeys":
    """Load a reference ``.ckpt`` / ``.pth`` (path or the loaded dict) into ``backbone`` the way main_linear.py does
    (``strict=False``).  Returns torch's missing / unexpected key report so callers can assert it is empty."""
    if isinstance(ckpt, str):
        assert ckpt.endswith((".ckpt", ".pth", ".pt"))
        ckpt = torch.load(ckpt, map_location="cpu")
    state = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    res = backbone.load_state_dict(rewrite_backbone_keys(state), strict=False)
    arena = getattr(backbone, "_arena", None)
    if arena is not None:
        arena.mark_dirty()
    return res
