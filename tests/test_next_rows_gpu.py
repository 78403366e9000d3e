"""GPU parity of the SURVEY.md §8(f) rows against the reference's golden outputs / the oracle: get_last_selfattention,
per-parameter gradient clipping, the LARS step (kernel level and through the fused engine), and the staged
(copy-stream) input path."""
import os

import numpy as np
import pytest
import torch

from oracle import chada_oracle as O
from oracle import det
from tests.golden import make_golden_f as MG
from tests.helpers import GOLDEN_DIR, cases, det_params, rel_err
from tests.test_next_rows_cpu import run_oracle_lars

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN_DIR, "reference_outputs_f.npz"))


@pytest.mark.parametrize("name", list(MG.ATTN_CASES))
def test_get_last_selfattention_vs_reference(name):
    from chadavit_b200.backbones import chada_vit
    c = MG.ATTN_CASES[name]
    m = chada_vit(patch_size=16, embed_dim=c["D"], return_all_tokens=False, max_number_channels=10)
    m.load_state_dict(det_params(O.backbone_shapes(c["D"]), c["seed"]))
    m = m.cuda()
    x = torch.from_numpy(det.det_pixels(c["n"], c["hw"], c["hw"], c["seed"])).cuda()
    A = m.get_last_selfattention(x)
    torch.cuda.synchronize()
    assert list(A.shape) == G[f"attn.{name}.shape"].tolist() and A.dtype == torch.float32
    ref = torch.from_numpy(G[f"attn.{name}.rows"])
    got = A[:, :, MG.ATTN_ROWS, :].cpu()
    err, rel = (got - ref).abs().max().item(), rel_err(got, ref)
    print(f"attention map {name}: max abs err {err:.2e} (max prob {ref.max().item():.3f}), rel L2 {rel:.2e}")
    assert err < 2.5e-2 * ref.max().item() and rel < 4e-2  # bf16 q/k/u operands: ~1e-2 relative on the probabilities
    assert (A.sum(-1) - 1).abs().max().item() < 1e-4
    # what main_attn.py:202-207 reads
    cls_map = A[0, :, 0, 1:].reshape(A.shape[1], -1)
    assert cls_map.shape == (2, (c["hw"] // 16) ** 2)


def test_get_last_selfattention_moyen_h12_vs_oracle():
    """Bare constructor: 12 heads of 16 (notebook path), D = 192."""
    from chadavit_b200.backbones import ChAdaViT
    P = det_params(O.backbone_shapes(192), 7)
    m = ChAdaViT(patch_size=16, embed_dim=192, return_all_tokens=False, max_number_channels=10)
    m.load_state_dict(P)
    m = m.cuda()
    x = torch.from_numpy(det.det_pixels(2, 224, 224, 8))
    A = m.get_last_selfattention(x.cuda()).cpu()
    with torch.no_grad():
        ref = O.last_selfattention(x, P, nhead=12)
    assert A.shape == ref.shape == (2, 12, 197, 197)
    err, rel = (A - ref).abs().max().item(), rel_err(A, ref)
    print(f"attention map moyen/h12: max abs err {err:.2e}, rel L2 {rel:.2e}")
    assert err < 1e-2 and rel < 6e-2


class _Bag(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])


def _arena_with(tensors):
    from chadavit_b200.arena import ParamArena
    bag = _Bag(tensors).cuda()
    return bag, ParamArena(bag)


def test_param_norms_and_clip_kernels_vs_reference():
    from chadavit_b200 import ops
    bag, ar = _arena_with(MG.lars_params())
    grads = [g * (10.0 if i % 2 == 0 else 0.01) for i, g in enumerate(MG.lars_inputs(0))]
    gflat = torch.zeros_like(ar.fp32)
    for n, g in zip(ar.names, grads):
        ar.g32(n, gflat).copy_(g.cuda() * 4.0)                       # the kernels see 4x the gradient and grad_scale = 1/4
    start, seg_of = ar.segment_maps()
    assert start.tolist() == [ar.offsets[n][0] // 64 for n in ar.names] + [ar.numel // 64]
    partial = torch.empty(ar.numel // 32, device="cuda")
    norms = torch.empty(3 * len(ar.names), device="cuda")
    seg_clip = torch.ones(len(ar.names), dtype=torch.uint8, device="cuda")
    ops.param_norms(ar.fp32, gflat, start, seg_clip, partial, norms, grad_scale=0.25, clip=0.3)
    nr = norms.cpu().view(-1, 3)
    for i, (p, g) in enumerate(zip(MG.lars_params(), grads)):
        coef = min(1.0, 0.3 / (g.norm().item() + 1e-6))
        assert abs(nr[i, 0].item() - p.norm().item()) < 1e-5 * p.norm().item()
        assert abs(nr[i, 2].item() - coef) < 1e-5 and abs(nr[i, 1].item() - g.norm().item() * coef) < 1e-5 * g.norm().item()
    ops.scale_grads(gflat, seg_of, norms)
    torch.cuda.synchronize()
    for i, n in enumerate(ar.names):
        ref = torch.from_numpy(G[f"clip.g{i}"])
        assert (ar.g32(n, gflat).cpu() * 0.25 - ref).abs().max().item() < 1e-6
    # reproducible bit for bit (no atomics): data-parallel replicas must stay identical
    norms2 = torch.empty_like(norms)
    ops.param_norms(ar.fp32, gflat, start, seg_clip, partial, norms2, grad_scale=0.25, clip=0.3)
    ops.param_norms(ar.fp32, gflat, start, seg_clip, partial, norms, grad_scale=0.25, clip=0.3)
    assert torch.equal(norms, norms2)


@pytest.mark.parametrize("name", list(MG.LARS_CONFIGS))
def test_lars_kernel_vs_reference(name):
    """Three LARS steps over a flat arena (parameter 4 receives its first gradient at step 1) against the reference's LARS."""
    from chadavit_b200 import ops
    c = dict(MG.LARS_CONFIGS[name])
    no_decay_1d = c.pop("no_decay_1d")
    bag, ar = _arena_with(MG.lars_params())
    teacher = ar.fp32.clone()
    t_ref = [p.clone() for p in MG.lars_params()]
    start, seg_of = ar.segment_maps()
    partial = torch.empty(ar.numel // 32, device="cuda")
    norms = torch.empty(3 * len(ar.names), device="cuda")
    buf = torch.zeros_like(ar.fp32)
    p16 = torch.empty(ar.numel, device="cuda", dtype=torch.bfloat16)
    stepped = set()
    # the reference trajectory, step by step, for the teacher EMA check
    c2 = dict(c); wd = c2.pop("weight_decay"); lr = c2.pop("lr")
    op, ob = MG.lars_params(), [None] * 5
    wds = [0.0 if (no_decay_1d and p.ndim <= 1) else wd for p in op]
    for step in range(3):
        gflat = torch.zeros_like(ar.fp32)
        flags = torch.full((ar.numel,), 2, dtype=torch.uint8)
        grads = MG.lars_inputs(step)
        for i, (n, p) in enumerate(zip(ar.names, ar.params)):
            off, cnt, _ = ar.offsets[n]
            if step == 0 and i == 4:
                continue                                              # no gradient: stays frozen this step
            ar.g32(n, gflat).copy_(grads[i].cuda())
            f = 0 if (no_decay_1d and p.dim() <= 1) else 1
            if p.dim() != 1 or not c.get("exclude_bias_n_norm", False):
                f |= 4
            if n not in stepped:
                f |= 8
                stepped.add(n)
            flags[off:off + cnt] = f
        ops.param_norms(ar.fp32, gflat, start, None, partial, norms)
        ops.lars_step(ar.fp32, gflat, buf, flags.cuda(), seg_of, norms, lr=c["lr"], momentum=c.get("momentum", 0.0),
                      dampening=c.get("dampening", 0.0), nesterov=c.get("nesterov", False), weight_decay=c["weight_decay"],
                      eta=c["eta"], clip_lr=c["clip_lr"], p_bf16=p16, teacher=teacher, tau=0.9)
        og = [None if (step == 0 and i == 4) else g for i, g in enumerate(grads)]
        op, ob = O.lars_step(op, og, ob, lr=lr, weight_decays=wds, **c2)
        t_ref = [0.9 * t + 0.1 * p for t, p in zip(t_ref, op)]
    torch.cuda.synchronize()
    for i, n in enumerate(ar.names):
        ref = torch.from_numpy(G[f"lars.{name}.p{i}"])
        got = ar.v32(n).cpu()
        assert (got - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item()), (name, i)
        off, cnt, shape = ar.offsets[n]
        assert (teacher[off:off + cnt].view(shape).cpu() - t_ref[i]).abs().max().item() < 2e-6
    assert torch.equal(p16, ar.fp32.to(torch.bfloat16))


def _dino(K, opt=None, clip=0.0, small=2, graph=False):
    from chadavit_b200.methods import DINO
    cfg = {"method": "dino", "backbone": {"kwargs": {"patch_size": 16, "embed_dim": 32, "return_all_tokens": False}},
           "data": {"max_img_channels": 10, "num_large_crops": 2, "num_small_crops": small},
           "method_kwargs": {"num_prototypes": K, "teacher_temperature": 0.07, "warmup_teacher_temperature_epochs": 0, "clip_grad": clip},
           "max_epochs": 10, "optimizer": opt or {"lr": 1e-3, "weight_decay": 0.01}, "engine": {"cuda_graph": graph}}
    return DINO(cfg)


def _step_fixture():
    st = cases()["step"]
    counts, K, sd = st["counts"], st["K"], st["seeds"]
    P = {"stu": det_params(O.backbone_shapes(32), sd["stu"]), "tea": det_params(O.backbone_shapes(32), sd["tea"]),
         "sh": det_params(O.head_shapes(32, K), sd["sh"]), "th": det_params(O.head_shapes(32, K), sd["th"])}
    crops = [torch.from_numpy(det.det_pixels(sum(counts), 224, 224, s_)) for s_ in sd["g"]] + \
            [torch.from_numpy(det.det_pixels(sum(counts), 96, 96, s_)) for s_ in sd["l"]]
    return counts, K, P, crops


def _load(m, P, epoch=1):
    m.backbone.load_state_dict(P["stu"]); m.momentum_backbone.load_state_dict(P["tea"])
    m.head.load_state_dict(P["sh"]); m.momentum_head.load_state_dict(P["th"])
    m = m.cuda()
    m.current_epoch = epoch
    m.on_train_epoch_start()
    return m


@pytest.mark.parametrize("epoch", [0, 1])
def test_engine_lars_step_with_clip(epoch):
    """fused_train_step with the pre-training yaml's optimizer (LARS, clip_lr, eta 0.02, exclude_bias_n_norm, momentum 0.9)
    and clip_grad: == gradients of the autograd path -> dino_clip_gradients -> freeze last layer -> LARS.step -> EMA, with
    clip and LARS taken from the oracle (pinned to the reference by tests/golden).  epoch 0: last layer frozen."""
    counts, K, P, crops = _step_fixture()
    opt = {"name": "lars", "lr": 0.3, "weight_decay": 1e-6, "exclude_bias_n_norm_wd": True,
           "kwargs": {"clip_lr": True, "eta": 0.02, "exclude_bias_n_norm": True, "momentum": 0.9}}
    clip = 0.02
    auto, fused = _load(_dino(K, opt, clip), P, epoch), _load(_dino(K, opt, clip), P, epoch)
    batch = ([c.cuda() for c in crops], None, [counts] * 4)
    loss = auto.training_step(batch)
    loss.backward()
    raw = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in auto.backbone.named_parameters()}
    auto.on_after_backward()                                          # product clip (kernels) + last-layer freeze
    torch.cuda.synchronize()
    # product clip == oracle clip on the same raw gradients
    names = [k for k, _ in auto.backbone.named_parameters()]
    clipped = O.clip_gradients([raw[k].cpu() if raw[k] is not None else None for k in names], clip)
    n_clipped = 0
    for k, ref, p in zip(names, clipped, auto.backbone.parameters()):
        if ref is None:
            continue
        assert (p.grad.cpu() - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item()), k
        n_clipped += int(not torch.equal(ref, raw[k].cpu()))
    assert 0 < n_clipped < len(names), n_clipped
    # reference update on those gradients (oracle LARS), two networks
    expect = {}
    for tag, mod in (("backbone", auto.backbone), ("head", auto.head)):
        ps = [p.detach().cpu() for p in mod.parameters()]
        gs = [p.grad.cpu() if p.grad is not None else None for p in mod.parameters()]
        wds = [0.0 if p.ndim <= 1 else 1e-6 for p in ps]
        newp, _ = O.lars_step(ps, gs, [None] * len(ps), lr=0.3, weight_decays=wds, momentum=0.9, eta=0.02, clip_lr=True, exclude_bias_n_norm=True)
        expect[tag] = newp
    tau = fused.momentum_updater.cur_tau
    loss_f = fused.fused_train_step(batch)
    torch.cuda.synchronize()
    assert abs(loss_f.item() - loss.item()) < 1e-5
    moved = 0
    for tag, mod, tmod, told in (("backbone", fused.backbone, fused.momentum_backbone, P["tea"]), ("head", fused.head, fused.momentum_head, P["th"])):
        for (k, p), ref, (_, t) in zip(mod.named_parameters(), expect[tag], tmod.named_parameters()):
            # split-K fp32 atomics make the two gradient computations differ in the last bits; LARS rescales by ||p||/||g||
            tol = 2e-3 * max(1e-3, (ref - P["stu" if tag == "backbone" else "sh"][k]).abs().max().item()) + 1e-6
            assert (p.detach().cpu() - ref).abs().max().item() <= tol, (tag, k)
            assert (t.detach().cpu() - (tau * told[k] + (1 - tau) * ref)).abs().max().item() <= 1e-5, (tag, k)
            moved += int(not torch.equal(ref, P["stu" if tag == "backbone" else "sh"][k]))
    lastv = fused.head.last_layer.weight_v.detach().cpu()
    assert torch.equal(lastv, P["sh"]["last_layer.weight_v"]) == (epoch == 0)      # frozen while epoch < freeze_last_layer
    assert torch.equal(fused.head.last_layer.weight_g.detach().cpu(), P["sh"]["last_layer.weight_g"])
    assert moved > 100


def test_engine_adamw_with_clip_matches_autograd_path():
    counts, K, P, crops = _step_fixture()
    auto, fused = _load(_dino(K, clip=0.02), P), _load(_dino(K, clip=0.02), P)
    batch = ([c.cuda() for c in crops], None, [counts] * 4)
    loss = auto.training_step(batch)
    loss.backward()
    opt = auto.configure_optimizers()
    auto.on_after_backward()
    opt.step()
    auto.on_train_batch_end()
    fused.fused_train_step(batch)
    torch.cuda.synchronize()
    for (k, p), (_, q) in zip(auto.named_parameters(), fused.named_parameters()):
        d = (p.detach() - q.detach()).abs().max().item()
        assert d <= 2e-3 * max(1.0, p.detach().abs().max().item()), (k, d)
    # and clipping did change the step
    plain = _load(_dino(K), P)
    plain.fused_train_step(batch)
    diff = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(plain.backbone.parameters(), fused.backbone.parameters()))
    assert diff > 1e-5


@pytest.mark.parametrize("graph", [False, True])
def test_staged_batches_equal_resident_batches(graph):
    """stage_batch (pinned host crops -> copy stream -> two alternating device buffer sets) feeds the same numbers as crops
    already on the device, over several steps with a different batch each step (eager and CUDA-graph replay)."""
    from chadavit_b200.data import OneChannelCollator
    counts, K, P, crops = _step_fixture()
    # lr = 0: the student never moves (split-K fp32 atomics make gradients differ in the last bits from run to run, and
    # Adam's first steps turn that into +-lr), so every forward is deterministic and the losses must agree to fp32 rounding;
    # teacher EMA, centre and temperature schedule still advance, and every step sees a different batch.
    zero = {"lr": 0.0, "weight_decay": 0.0}
    a, b = _load(_dino(K, opt=zero, graph=graph), P), _load(_dino(K, opt=zero, graph=graph), P)
    coll = OneChannelCollator(pin_memory=True)
    off = np.concatenate([[0], np.cumsum(counts)])

    def host_batch(step):
        samples = []
        for i, C in enumerate(counts):
            imgs = [(c[off[i]:off[i + 1], 0] * (1.0 + 0.25 * step)).contiguous() for c in crops]
            samples.append((i, imgs, i))
        return coll(samples)

    nxt = b.stage_batch(host_batch(0))
    losses = []
    for step in range(5):
        hb = host_batch(step)
        assert hb[2] == [counts] * 4 and all(c.is_pinned() for c in hb[0])
        la = a.fused_train_step(([c.cuda() for c in hb[0]], None, hb[2]))
        cur, nxt = nxt, None
        lb = b.fused_train_step(cur)
        if step + 1 < 5:
            nxt = b.stage_batch(host_batch(step + 1))          # overlaps the step just launched
        assert abs(la.item() - lb.item()) < 2e-6, (step, la.item(), lb.item())
        for dcrop, hcrop in zip(cur[0], hb[0]):                # the staged device buffers hold exactly this step's crops
            assert torch.equal(dcrop.cpu(), hcrop)
        losses.append(la.item())
    assert len(set(losses)) == 5                               # every step saw a different batch
    torch.cuda.synchronize()
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p.detach(), q.detach()), k
    assert torch.equal(a.dino_loss_func.center, b.dino_loss_func.center)


def test_non_square_images_vs_reference():
    """96 x 224 images (6 x 14 patches): packed tokenizer + bicubic position-grid resize against the reference's output."""
    from chadavit_b200.backbones import chada_vit
    c = MG.NONSQ
    m = chada_vit(patch_size=16, embed_dim=c["D"], return_all_tokens=False, max_number_channels=10)
    m.load_state_dict(det_params(O.backbone_shapes(c["D"]), c["seed"]))
    m = m.cuda()
    x = torch.from_numpy(det.det_pixels(sum(c["counts"]), c["H"], c["W"], c["seed"])).cuda()
    with torch.no_grad():
        y = m(x, 0, [c["counts"]]).cpu()
    ref = torch.from_numpy(G["nonsq.out"])
    assert y.shape == ref.shape and rel_err(y, ref) < 1e-2
