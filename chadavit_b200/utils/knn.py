"""Weighted k-NN classifier behind the reference's interface (src/utils/knn.py:27-177; used by main_knn.py:99 and
base.py:285).  Same constructor, ``update`` / ``compute`` / ``reset`` and the same voting rule; the feature-bank similarity
matrix — the only heavy part — runs on the tcgen05 GEMM with split-bf16 operands (fp32-accurate, see csrc/knn.cu), row
normalisation fused into the operand preparation.  Top-k selection and the k x classes vote are small torch index ops.
There is no torchmetrics dependency: distributed runs gather features with ``torch.distributed`` before ``compute``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import ops


class WeightedKNNClassifier:
    def __init__(self, k: int = 20, T: float = 0.07, max_distance_matrix_size: int = int(5e6), distance_fx: str = "cosine",
                 epsilon: float = 0.00001, dist_sync_on_step: bool = False):
        self.k = k
        self.T = T
        self.max_distance_matrix_size = max_distance_matrix_size
        self.distance_fx = distance_fx
        self.epsilon = epsilon
        self.dist_sync_on_step = dist_sync_on_step
        self.reset()

    def reset(self) -> None:
        self.train_features: List[torch.Tensor] = []
        self.train_targets: List[torch.Tensor] = []
        self.test_features: List[torch.Tensor] = []
        self.test_targets: List[torch.Tensor] = []

    def update(self, train_features: Optional[torch.Tensor] = None, train_targets: Optional[torch.Tensor] = None,
               test_features: Optional[torch.Tensor] = None, test_targets: Optional[torch.Tensor] = None) -> None:
        """Appends a batch to the train and / or test bank; features and targets of a bank come together (knn.py:62-94)."""
        for feats, targets, bank_f, bank_t in ((train_features, train_targets, self.train_features, self.train_targets),
                                               (test_features, test_targets, self.test_features, self.test_targets)):
            if (feats is None) != (targets is None):
                raise AssertionError("features and targets of a bank must be passed together")
            if feats is None:
                continue
            if feats.size(0) != targets.size(0):
                raise AssertionError(f"{feats.size(0)} feature rows for {targets.size(0)} targets")
            bank_f.append(feats.detach())
            bank_t.append(targets.detach())

    __call__ = update

    @torch.no_grad()
    def compute(self) -> Tuple[float, float]:
        """Weighted k-NN accuracy @1 and @5 (knn.py:96-177)."""
        if not self.train_features or not self.test_features:
            return -1, -1
        if self.distance_fx not in ("cosine", "euclidean"):
            raise NotImplementedError
        train_features = torch.cat(self.train_features).float().contiguous()
        train_targets = torch.cat(self.train_targets)
        test_features = torch.cat(self.test_features).float().contiguous()
        test_targets = torch.cat(self.test_targets)
        if not train_features.is_cuda:
            raise RuntimeError("chadavit_b200.WeightedKNNClassifier needs CUDA features (no CPU fallback)")
        cosine = self.distance_fx == "cosine"
        num_classes = torch.unique(test_targets).numel()
        num_train_images = train_targets.size(0)
        num_test_images = test_targets.size(0)
        chunk_size = min(max(1, self.max_distance_matrix_size // num_train_images), num_test_images)
        k = min(self.k, num_train_images)
        # the bank is prepared once: rows normalised (cosine), split into bf16 hi/lo, padded to the GEMM's N granularity
        bank, bank_sq = ops.split_bf16x3(train_features, role_b=True, normalize=cosine, want_sqnorm=not cosine, pad_rows_to=8)
        top1, top5, total = 0.0, 0.0, 0
        for idx in range(0, num_test_images, chunk_size):
            features = test_features[idx:min(idx + chunk_size, num_test_images)]
            targets = test_targets[idx:min(idx + chunk_size, num_test_images)]
            batch_size = targets.size(0)
            q, q_sq = ops.split_bf16x3(features, role_b=False, normalize=cosine, want_sqnorm=not cosine)
            sims = ops.gemm(q, bank, flags=ops.EPI_OUT_F32)                      # [batch, padded bank] fp32
            if not cosine:
                ops.inv_euclid_(sims, q_sq, bank_sq, num_train_images, self.epsilon)
            similarities, indices = sims[:, :num_train_images].topk(k, largest=True, sorted=True)
            retrieved_neighbors = train_targets[indices.reshape(-1)].view(batch_size, k)
            if cosine:
                similarities = similarities.div(self.T).exp_()
            probs = torch.zeros(batch_size, num_classes, device=sims.device, dtype=torch.float32)
            probs.scatter_add_(1, retrieved_neighbors.long(), similarities)      # sum of weights per class (knn.py:156-162)
            _, predictions = probs.sort(1, True)
            correct = predictions.eq(targets.view(-1, 1))
            top1 += correct.narrow(1, 0, 1).sum().item()
            top5 += correct.narrow(1, 0, min(5, k, correct.size(-1))).sum().item()
            total += batch_size
        self.reset()
        return top1 * 100.0 / total, top5 * 100.0 / total
