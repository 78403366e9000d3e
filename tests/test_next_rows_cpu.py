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
