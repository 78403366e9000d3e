"""Embedding extraction behind BaseMethod._base_extract_step (src/methods/base.py:927-981), multi_channels strategy:
the call main_knn.py / the UMAP and linear-probe drivers make to turn a collated ragged batch into one feature row per image."""
from __future__ import annotations

from typing import List

import torch

from ..backbones.chada_vit import ChAdaViT


@torch.no_grad()
def extract_features(backbone: ChAdaViT, X: torch.Tensor, index: int, list_num_channels: List[List[int]], *,
                     mixed_channels: bool = False, no_channel_last: bool = False) -> torch.Tensor:
    """``feats`` of _base_extract_step: CLS embeddings ``(B, D)`` or, for a backbone built with ``return_all_tokens``, the
    per-image concatenation of all patch tokens ``(B, C*N*D)`` (equal channel counts required, as in the reference's
    ``torch.stack``; ``mixed_channels`` returns the ragged ``(ΣC*N, D)`` token matrix untouched, base.py:958)."""
    assert isinstance(backbone, ChAdaViT), "Only backbone of class ChAdaViT is currently supported for multi_channels strategy."
    if not no_channel_last:
        X = X.to(memory_format=torch.channels_last)              # base.py:941-942: a no-op for (ΣC, 1, H, W) memory
    feats = backbone(X, index, list_num_channels)
    if mixed_channels or not backbone.return_all_tokens:
        return feats
    counts = list_num_channels[index]
    chunks = feats.view(sum(counts), -1, feats.shape[-1])        # (ΣC, N, D)
    chunks = torch.split(chunks, counts, dim=0)
    return torch.stack(chunks, dim=0).flatten(start_dim=1)
