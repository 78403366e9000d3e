"""CPU: SURVEY.md §8(f) rows — the oracle's restatements of get_last_selfattention, dino_clip_gradients, LARS.step and
one_channel_collate_fn reproduce the golden outputs of the REFERENCE (tests/golden/make_golden_f.py), and the product's
host-side collate is bit-identical to them."""
import os

import numpy as np
import pytest
import torch

from oracle import chada_oracle as O
from oracle import det
from tests.golden import make_golden_f as MG
from tests.helpers import GOLDEN_DIR

G = np.load(os.path.join(GOLDEN_DIR, "reference_outputs_f.npz"))


@pytest.mark.parametrize("name", list(MG.ATTN_CASES))
def test_oracle_last_selfattention(name):
    c = MG.ATTN_CASES[name]
    P = {k: torch.from_numpy(v) for k, v in det.det_state_dict(O.backbone_shapes(c["D"]), c["seed"]).items()}
    x = torch.from_numpy(det.det_pixels(c["n"], c["hw"], c["hw"], c["seed"]))
    with torch.no_grad():
        A = O.last_selfattention(x, P, nhead=2)
    assert list(A.shape) == G[f"attn.{name}.shape"].tolist()
    assert (A[:, :, MG.ATTN_ROWS, :] - torch.from_numpy(G[f"attn.{name}.rows"])).abs().max().item() < 5e-6
    assert (A.sum(-1) - 1).abs().max().item() < 1e-5


def run_oracle_lars(name):
    c = dict(MG.LARS_CONFIGS[name])
    no_decay_1d = c.pop("no_decay_1d")
    wd = c.pop("weight_decay")
    lr = c.pop("lr")
    params = MG.lars_params()
    wds = [0.0 if (no_decay_1d and p.ndim <= 1) else wd for p in params]
    bufs = [None] * len(params)
    for step in range(3):
        grads = [None if (step == 0 and i == 4) else g for i, g in enumerate(MG.lars_inputs(step))]
        params, bufs = O.lars_step(params, grads, bufs, lr=lr, weight_decays=wds, **c)
    return params


@pytest.mark.parametrize("name", list(MG.LARS_CONFIGS))
def test_oracle_lars(name):
    for i, p in enumerate(run_oracle_lars(name)):
        ref = torch.from_numpy(G[f"lars.{name}.p{i}"])
        assert (p - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item()), (name, i)


def test_oracle_clip_gradients():
    grads = [g * (10.0 if i % 2 == 0 else 0.01) for i, g in enumerate(MG.lars_inputs(0))]
    out = O.clip_gradients(grads, 0.3)
    changed = 0
    for i, g in enumerate(out):
        ref = torch.from_numpy(G[f"clip.g{i}"])
        assert (g - ref).abs().max().item() <= 1e-7
        changed += int(not torch.equal(g, grads[i]))
    assert 0 < changed < len(out)          # the fixture exercises both branches


@pytest.mark.parametrize("impl", ["oracle", "product", "product_pinned_reuse"])
def test_collate_bit_exact(impl):
    if impl == "oracle":
        fn = O.one_channel_collate
    else:
        from chadavit_b200.data import OneChannelCollator, one_channel_collate_fn
        fn = one_channel_collate_fn if impl == "product" else OneChannelCollator(pin_memory=False, reuse=2)
    for rep in range(3 if impl == "product_pinned_reuse" else 1):      # buffer reuse must not leak earlier contents
        crops, labels, counts = fn(MG.collate_batch())
        assert isinstance(crops, list) and len(crops) == 3
        for i, c in enumerate(crops):
            ref = G[f"collate.crop{i}"]
            assert tuple(c.shape) == ref.shape and c.dtype == torch.float32
            assert np.array_equal(c.numpy(), ref)                      # bit-exact
        assert labels.dtype == torch.int64 and labels.tolist() == G["collate.labels"].tolist()
        assert counts == G["collate.counts"].tolist() and all(isinstance(v, int) for l in counts for v in l)
    single = [(t[1][2], t[2]) for t in MG.collate_batch()]
    x1, _, c1 = fn(single)
    assert isinstance(x1, torch.Tensor) and np.array_equal(x1.numpy(), G["collate.single"])
    assert c1 == G["collate.single_counts"].tolist()


def test_collate_rejects_mixed_sizes():
    from chadavit_b200.data import one_channel_collate_fn
    bad = [(torch.zeros(2, 32, 32), 0), (torch.zeros(1, 16, 16), 1)]
    with pytest.raises(RuntimeError):
        one_channel_collate_fn(bad)
    with pytest.raises(RuntimeError):
        O.one_channel_collate(bad)


@pytest.mark.parametrize("name", list(MG.KNN_CASES))
def test_oracle_knn(name):
    xtr, ytr, xte, yte = MG.knn_data()
    t1, t5, _ = O.knn_compute(xtr, ytr, xte, yte, **MG.KNN_CASES[name])
    assert [t1, t5] == pytest.approx(G[f"knn.{name}"].tolist(), abs=1e-9)


def test_checkpoint_key_rewriting_and_load():
    """A Lightning-style DINO state_dict (backbone.*, momentum_backbone.*, head.* ...) loads into the backbone the way
    main_linear.py:103-110 does it: same surviving keys as the restated rule, nothing missing, values bit-equal."""
    from chadavit_b200.backbones import chada_vit
    from chadavit_b200.utils.checkpoint import load_pretrained_backbone, rewrite_backbone_keys
    P = {k: torch.from_numpy(v) for k, v in det.det_state_dict(O.backbone_shapes(32), 3).items()}
    H = {k: torch.from_numpy(v) for k, v in det.det_state_dict(O.head_shapes(32, 64), 4).items()}
    state = {f"backbone.{k}": v for k, v in P.items()}
    state.update({f"head.{k}": v for k, v in H.items()})
    state.update({f"momentum_head.{k}": v + 1 for k, v in H.items()})
    state["dino_loss_func.center"] = torch.zeros(1, 64)
    want = O.rewrite_checkpoint_keys(state)
    got = rewrite_backbone_keys(state)
    assert list(got.keys()) == list(want.keys()) and set(P.keys()) <= set(got.keys())
    m = chada_vit(patch_size=16, embed_dim=32, return_all_tokens=False, max_number_channels=10)
    res = load_pretrained_backbone(m, {"state_dict": state})
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, P[k]), k
    enc = {k.replace("backbone.", "encoder."): v for k, v in state.items() if k.startswith("backbone.")}
    # quirk kept on purpose: 'encoder.*' keys are renamed to 'backbone.*' AFTER the key snapshot was taken, so the prefix is never
    # stripped from them and strict=False then loads nothing (main_linear.py:103-110 behaves exactly like this)
    assert list(rewrite_backbone_keys(enc).keys()) == list(O.rewrite_checkpoint_keys(enc).keys()) == list(state.keys())[:len(enc)]
    res = load_pretrained_backbone(chada_vit(patch_size=16, embed_dim=32, return_all_tokens=False, max_number_channels=10), enc)
    assert len(res.missing_keys) == len(P) and len(res.unexpected_keys) == len(P)


def test_oracle_extract_features_shapes():
    toks = torch.arange(6 * 4 * 2, dtype=torch.float32).view(6 * 4, 2)        # ΣC = 6 channel images, N = 4, D = 2
    out = O.extract_features(toks, [3, 3], return_all_tokens=True)
    assert out.shape == (2, 3 * 4 * 2) and torch.equal(out[1], toks[12:].reshape(-1))
    assert O.extract_features(toks, [2, 4], return_all_tokens=True, mixed_channels=True) is toks
    with pytest.raises(RuntimeError):
        O.extract_features(toks, [2, 4], return_all_tokens=True)              # torch.stack of unequal chunks (base.py:975)


def test_knn_bank_bookkeeping_and_no_cpu_fallback():
    from chadavit_b200.utils.knn import WeightedKNNClassifier
    knn = WeightedKNNClassifier(k=3, T=0.1, distance_fx="cosine")
    assert knn.compute() == (-1, -1)                                   # knn.py:109-110
    knn(train_features=torch.zeros(4, 8), train_targets=torch.zeros(4, dtype=torch.long))
    knn.update(test_features=torch.zeros(2, 8), test_targets=torch.zeros(2, dtype=torch.long))
    assert len(knn.train_features) == len(knn.test_features) == 1
    with pytest.raises(AssertionError):
        knn.update(train_features=torch.zeros(1, 8))                   # features without targets (knn.py:78)
    with pytest.raises(AssertionError):
        knn.update(test_features=torch.zeros(3, 8), test_targets=torch.zeros(2))
    with pytest.raises(RuntimeError):
        knn.compute()                                                  # CPU features: the product path has no fallback


def test_engine_optimizer_flags_host_logic():
    """Per-element flag bytes the fused optimizers read: bit0 weight decay (base.py:426-427 via exclude_bias_n_norm_wd), bit1
    frozen (weight_g under norm_last_layer, alignment padding, last_layer while epoch < freeze_last_layer), bit2 LARS layer-wise
    adaptation (lars.py:136)."""
    from chadavit_b200.methods import DINO
    cfg = {"method": "dino", "backbone": {"kwargs": {"patch_size": 16, "embed_dim": 32, "return_all_tokens": False}},
           "data": {"max_img_channels": 10, "num_large_crops": 2, "num_small_crops": 0},
           "method_kwargs": {"num_prototypes": 64, "clip_grad": 3.0},
           "optimizer": {"name": "lars", "lr": 0.3, "weight_decay": 1e-6, "exclude_bias_n_norm_wd": True,
                         "kwargs": {"clip_lr": True, "eta": 0.02, "exclude_bias_n_norm": True}}}
    m = DINO(cfg)
    assert m.optimizer == "lars" and m.lars["momentum"] == 0.9 and m.clip_grad == 3.0       # src/args/pretrain.py:221 default
    for name, net in (("backbone", m.backbone), ("head", m.head)):
        ar = net.arena
        st = m._opt_state(name, ar)
        fl = st["flags"]
        assert "v" not in st and st["norms"].numel() == 3 * len(ar.names)                   # LARS: one momentum buffer, norms workspace
        used = torch.zeros(ar.numel, dtype=torch.bool)
        for n, p in zip(ar.names, ar.params):
            off, cnt, _ = ar.offsets[n]
            used[off:off + cnt] = True
            want = (0 if p.dim() <= 1 else 1) | (0 if p.requires_grad else 2) | (4 if p.dim() != 1 else 0)
            if name == "head" and n.startswith("last_layer."):
                want |= 16                                                                   # AdamW: the parameter's own step count
            assert (fl[off:off + cnt] == want).all(), (n, want, fl[off].item())
        assert (fl[~used] == 2).all()                                                        # padding never moves
        start, seg_of = ar.segment_maps()
        assert seg_of.numel() == ar.numel // 64 and start[-1].item() == ar.numel // 64
        assert all(seg_of[ar.offsets[n][0] // 64].item() == i for i, n in enumerate(ar.names))
    hd = m._opt_state("head", m.head.arena)
    off, cnt, _ = m.head.arena.offsets["last_layer.weight_v"]
    assert (hd["flags_frozen_last"][off:off + cnt] == 2).all() and (hd["flags"][off:off + cnt] & 2 == 0).all()
    assert m.head.last_layer.weight_g.requires_grad is False
    with pytest.raises(NotImplementedError):
        DINO({**cfg, "optimizer": {"name": "sgd"}})
    from chadavit_b200.utils.lars import LARS
    opt = m.configure_optimizers()                                                           # the reference's LARS interface (lars.py:21-111)
    assert isinstance(opt, LARS) and isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 4
    assert [g["weight_decay"] for g in opt.param_groups] == [1e-6, 0.0, 1e-6, 0.0] and opt.param_groups[0]["eta"] == 0.02
    assert all(p.requires_grad for g in opt.param_groups for p in g["params"])
    foreign = torch.nn.Parameter(torch.zeros(3))
    foreign.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        LARS([foreign], lr=0.1).step()                                                       # bare tensors: no per-tensor fallback


def test_non_square_images_oracle_and_position_grid():
    """A 96 x 224 image (6 x 14 patches): the oracle reproduces the reference's output, and the product's precomputed linear map
    of the bicubic position-grid resize (ChAdaViT._interp_matrix, host code) equals the reference interpolation."""
    from chadavit_b200.backbones import chada_vit
    c = MG.NONSQ
    P = {k: torch.from_numpy(v) for k, v in det.det_state_dict(O.backbone_shapes(c["D"]), c["seed"]).items()}
    x = torch.from_numpy(det.det_pixels(sum(c["counts"]), c["H"], c["W"], c["seed"]))
    with torch.no_grad():
        y = O.backbone_forward(x, 0, [c["counts"]], P, nhead=2, final_eps=1e-6)
    assert (y - torch.from_numpy(G["nonsq.out"])).abs().max().item() < 2e-5
    m = chada_vit(patch_size=16, embed_dim=c["D"], return_all_tokens=False, max_number_channels=10)
    m.load_state_dict(P)
    for (H, W) in ((96, 224), (96, 96), (224, 96)):
        hp, wp = H // 16, W // 16
        M = m._interp_matrix(hp, wp, H, W, "cpu")
        assert M.shape == (hp * wp, 196)
        ref = O.interp_pos_embed(P["pos_embed"], hp * wp, H, W, 16)[0, 0]
        assert (M @ P["pos_embed"][0, 0, 1:] - ref).abs().max().item() < 1e-6


def test_warmup_cosine_lr_matches_reference_scheduler():
    """Closed-form schedule for the engine step == the reference's LinearWarmupCosineAnnealingLR stepped once per update
    (src/utils/lr_scheduler.py:14-150; skipped on machines without the reference tree: the formula is also checked at its
    fixed points)."""
    import importlib.util
    from chadavit_b200.utils.lr_schedule import warmup_cosine_lr
    from oracle import ref_loader
    kw = dict(base_lr=0.3, warmup_steps=10, max_steps=60, warmup_start_lr=3e-5, eta_min=1e-4)
    assert warmup_cosine_lr(0, **kw) == 3e-5 and abs(warmup_cosine_lr(9, **kw) - 0.3) < 1e-12
    assert abs(warmup_cosine_lr(10, **kw) - 0.3) < 1e-12 and abs(warmup_cosine_lr(60, **kw) - 1e-4) < 1e-12
    assert abs(warmup_cosine_lr(35, **kw) - (1e-4 + 0.5 * (0.3 - 1e-4))) < 1e-12          # half way down the cosine
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("_ref_lr_sched", os.path.join(ref_loader.REF_ROOT, "src/utils/lr_scheduler.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=kw["base_lr"])
    sched = mod.LinearWarmupCosineAnnealingLR(opt, warmup_epochs=kw["warmup_steps"], max_epochs=kw["max_steps"],
                                              warmup_start_lr=kw["warmup_start_lr"], eta_min=kw["eta_min"])
    for step in range(60):
        assert abs(opt.param_groups[0]["lr"] - warmup_cosine_lr(step, **kw)) < 1e-9, step
        opt.step()
        sched.step()
