"""GPU: BASELINE.json's full-size configurations through size-independent properties (the CPU oracle cannot run them in
seconds): per-image independence of a ragged batch (== the reference's key-padding semantics: an image never sees another
image's or a padded channel's tokens), permutation equivariance, linearity of the backward pass in the output gradient."""
import numpy as np
import pytest
import torch

from oracle import chada_oracle as O
from oracle import det
from tests.helpers import det_params, rel_err

pytestmark = pytest.mark.gpu


def test_config1_moyen_embedding_extraction_batch_256():
    """configs[1]: ChAda-ViT-moyen/16, ragged 1-10 channel batch of 256, CLS embeddings."""
    from chadavit_b200.backbones import chada_vit
    from chadavit_b200.methods import extract_features
    m = chada_vit(patch_size=16, embed_dim=192, return_all_tokens=False, max_number_channels=10)
    m.load_state_dict(det_params(O.backbone_shapes(192), 5))
    m = m.cuda()
    counts = np.random.RandomState(1234).randint(1, 11, size=256).tolist()          # HOW_TO_USE.ipynb cell 16
    off = np.concatenate([[0], np.cumsum(counts)])
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(int(off[-1]), 1, 224, 224, device="cuda", generator=g)
    with torch.no_grad():
        full = extract_features(m, x, 0, [counts])
        assert full.shape == (256, 192) and torch.isfinite(full).all()
        # (a) independence: an image embedded alone or in a small batch == the same image inside the batch of 256
        for sel in ([0], [17, 200], [255, 3, 128]):
            xs = torch.cat([x[off[b]:off[b + 1]] for b in sel])
            sub = m(xs, 0, [[counts[b] for b in sel]])
            d = (sub - full[sel]).abs().max().item()
            assert d <= 1e-5, (sel, d)
        # (b) permutation equivariance over images
        perm = np.random.RandomState(1).permutation(256)
        xp = torch.cat([x[off[b]:off[b + 1]] for b in perm])
        outp = m(xp, 0, [[counts[b] for b in perm]])
        assert (outp - full[perm]).abs().max().item() <= 1e-5
        # (c) the index argument selects the channel list (chada_vit.py:226)
        again = m(x, 1, [[1], counts])
        assert torch.equal(again, full)


def test_config4_base_all_ten_channels_fwd_bwd():
    """configs[4]: ChAda-ViT-base/16 (D = 768, 12 heads of 64), every image with 10 channels (1961-token sequences)."""
    from chadavit_b200.backbones import ChAdaViT
    m = ChAdaViT(patch_size=16, embed_dim=768, num_heads=12, return_all_tokens=False, max_number_channels=10)
    m.load_state_dict(det_params(O.backbone_shapes(768), 9))
    m = m.cuda()
    counts = [10, 10, 10]
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(30, 1, 224, 224, device="cuda", generator=g)
    w1 = torch.randn(3, 768, device="cuda", generator=g)
    w2 = torch.randn(3, 768, device="cuda", generator=g)

    def grads(w):
        for p in m.parameters():
            p.grad = None
        y = m(x, 0, [counts])
        (y * w).sum().backward()
        return y.detach(), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}

    y, ga = grads(w1)
    _, gb = grads(w2)
    _, gab = grads(w1 + w2)
    assert y.shape == (3, 768) and torch.isfinite(y).all()
    with torch.no_grad():
        alone = m(x[10:20], 0, [[10]])
    assert (alone - y[1:2]).abs().max().item() <= 1e-5                                   # independence at S = 1961
    # backward is linear in the output gradient (up to bf16 rounding of the intermediate gradients)
    worst = 0.0
    for k in ("cls_token", "channel_token", "pos_embed", "blocks.0.self_attn.in_proj_weight", "blocks.5.linear1.weight",
              "blocks.11.linear2.weight", "blocks.7.norm1.weight", "token_learner.proj.weight", "norm.bias"):
        e = rel_err(gab[k], ga[k] + gb[k])
        worst = max(worst, e)
        assert torch.isfinite(gab[k]).all() and e < 3e-2, (k, e)
    print(f"base/16 all-10-channel: backward linearity worst rel err {worst:.2e}")
    assert float(ga["channel_token"].abs().sum()) > 0


def test_all_token_extraction_matches_oracle():
    """_base_extract_step with return_all_tokens (base.py:966-979): one row of C*N*D features per image."""
    from chadavit_b200.backbones import chada_vit
    from chadavit_b200.methods import extract_features
    P = det_params(O.backbone_shapes(32), 2)
    m = chada_vit(patch_size=16, embed_dim=32, return_all_tokens=True, max_number_channels=10)
    m.load_state_dict(P)
    m = m.cuda()
    counts = [3, 3]
    x = torch.from_numpy(det.det_pixels(6, 96, 96, 2))
    got = extract_features(m, x.cuda(), 0, [counts]).cpu()
    with torch.no_grad():
        toks = O.backbone_forward(x, 0, [counts], P, nhead=2, final_eps=1e-6, return_all_tokens=True)
        ref = O.extract_features(toks, counts, return_all_tokens=True)
    assert got.shape == ref.shape == (2, 3 * 36 * 32)
    assert rel_err(got, ref) < 1e-2
    ragged = extract_features(m, x.cuda(), 0, [[2, 4]], mixed_channels=True)
    assert ragged.shape == (6 * 36, 32)


@pytest.mark.parametrize("name", ["cosine", "cosine_k200", "euclidean"])
def test_weighted_knn_matches_reference(name):
    import os
    from chadavit_b200.utils.knn import WeightedKNNClassifier
    from tests.golden import make_golden_f as MG
    from tests.helpers import GOLDEN_DIR
    G = np.load(os.path.join(GOLDEN_DIR, "reference_outputs_f.npz"))
    xtr, ytr, xte, yte = MG.knn_data()
    knn = WeightedKNNClassifier(**MG.KNN_CASES[name])
    knn(train_features=xtr[:100].cuda(), train_targets=ytr[:100].cuda())
    knn.update(train_features=xtr[100:].cuda(), train_targets=ytr[100:].cuda(), test_features=xte.cuda(), test_targets=yte.cuda())
    t1, t5 = knn.compute()
    r1, r5 = G[f"knn.{name}"].tolist()
    print(f"knn {name}: top1 {t1:.3f} (reference {r1:.3f}) top5 {t5:.3f} (reference {r5:.3f})")
    assert abs(t1 - r1) < 1e-6 and abs(t5 - r5) < 1e-6
    assert knn.compute() == (-1, -1)                           # compute() resets the banks (knn.py:175)
    # the similarity matrix itself: fp32-grade on the bf16 tensor cores
    from chadavit_b200 import ops
    a, _ = ops.split_bf16x3(xte.cuda().contiguous(), role_b=False, normalize=True)
    b, _ = ops.split_bf16x3(xtr.cuda().contiguous(), role_b=True, normalize=True, pad_rows_to=8)
    sim = ops.gemm(a, b, flags=ops.EPI_OUT_F32)[:, :xtr.shape[0]].cpu()
    ref = torch.nn.functional.normalize(xte) @ torch.nn.functional.normalize(xtr).t()
    err = (sim - ref).abs().max().item()
    print(f"split-bf16 similarity: max abs err {err:.2e}")
    assert err < 5e-5


def test_moyen_lars_clip_graph_steps():
    """configs[2] shape (moyen/16, 2 global + 6 local crops) at a small batch with the pre-training yaml's optimizer (LARS,
    clip_lr, exclude_bias_n_norm) + clip_grad, through CUDA-graph replay: finite, parameters and teacher move, the frozen
    prototypes do not (epoch 0 < freeze_last_layer), weight_g stays 1."""
    from chadavit_b200.methods import DINO
    cfg = {"method": "dino", "backbone": {"kwargs": {"patch_size": 16, "embed_dim": 192, "return_all_tokens": False}},
           "data": {"max_img_channels": 10, "num_large_crops": 2, "num_small_crops": 6},
           "method_kwargs": {"num_prototypes": 4096, "clip_grad": 3.0, "freeze_last_layer": 1},
           "max_epochs": 100, "max_steps": 1000, "momentum": {"base_tau": 0.99, "final_tau": 1.0},
           "optimizer": {"name": "lars", "lr": 0.3, "weight_decay": 1e-6, "exclude_bias_n_norm_wd": True,
                         "kwargs": {"clip_lr": True, "eta": 0.02, "exclude_bias_n_norm": True, "momentum": 0.9}},
           "engine": {"cuda_graph": True}}
    torch.manual_seed(0)
    m = DINO(cfg).cuda()
    counts = [3, 1, 10, 5, 2, 7]
    g = torch.Generator(device="cuda").manual_seed(3)
    crops = [torch.randn(sum(counts), 1, 224, 224, device="cuda", generator=g) for _ in range(2)] + \
            [torch.randn(sum(counts), 1, 96, 96, device="cuda", generator=g) for _ in range(6)]
    batch = (crops, None, [counts] * 8)
    p0 = m.backbone.arena.fp32.clone() if m.backbone._arena is not None else None
    w0 = {k: v.detach().clone() for k, v in m.backbone.state_dict().items()}
    t0 = {k: v.detach().clone() for k, v in m.momentum_backbone.state_dict().items()}
    v0 = m.head.last_layer.weight_v.detach().clone()
    losses = [m.fused_train_step(batch).item() for _ in range(5)]        # eager, capture, 3 replays
    torch.cuda.synchronize()
    assert all(np.isfinite(l) for l in losses), losses
    assert m.use_cuda_graph, "graph capture fell back to eager"
    moved = sum(int(not torch.equal(v, w0[k])) for k, v in m.backbone.state_dict().items())
    assert moved == len(w0), (moved, len(w0))
    assert all(torch.isfinite(v).all() for v in m.backbone.state_dict().values())
    assert any(not torch.equal(v, t0[k]) for k, v in m.momentum_backbone.state_dict().items())
    assert torch.equal(m.head.last_layer.weight_v.detach(), v0)           # frozen in epoch 0
    assert torch.equal(m.head.last_layer.weight_g.detach(), torch.ones_like(m.head.last_layer.weight_g))
    # clip_lr bounds every LARS update: |dp| <= lr * |momentum-averaged (g + wd p)| with lars_lr <= 1
    step = max((v - w0[k]).abs().max().item() for k, v in m.backbone.state_dict().items())
    assert step < 10.0, step
