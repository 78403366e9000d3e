"""GPU: round-2 engine features — AdamW across the last-layer freeze boundary, the reference's LARS interface on the autograd
path, CUDA graphs surviving layout-cache eviction, engine checkpoint round trip, and (2 GPUs) W-rank data-parallel steps ==
one process on the concatenated batch (SURVEY.md §8e equivalence test)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import chada_oracle as O
from oracle import det
from tests.helpers import det_params, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(K=256, opt=None, graph=False, small=1, freeze=1, D=32, extra_engine=None):
    return {"method": "dino", "backbone": {"kwargs": {"patch_size": 16, "embed_dim": D, "return_all_tokens": False}},
            "data": {"max_img_channels": 10, "num_large_crops": 2, "num_small_crops": small},
            "method_kwargs": {"num_prototypes": K, "freeze_last_layer": freeze, "warmup_teacher_temperature_epochs": 0},
            "max_epochs": 10, "max_steps": 1000, "optimizer": opt or {"lr": 1e-3, "weight_decay": 0.01},
            "engine": {"cuda_graph": graph, **(extra_engine or {})}}


def _build(cfg, salt=0, D=32):
    from chadavit_b200.methods import DINO
    K = cfg["method_kwargs"]["num_prototypes"]
    m = DINO(cfg)
    m.backbone.load_state_dict(det_params(O.backbone_shapes(D), 61 + salt)); m.momentum_backbone.load_state_dict(det_params(O.backbone_shapes(D), 62 + salt))
    m.head.load_state_dict(det_params(O.head_shapes(D, K), 63 + salt)); m.momentum_head.load_state_dict(det_params(O.head_shapes(D, K), 64 + salt))
    return m.cuda()


def _batch(counts, seed, small=1, dev="cuda"):
    G = sum(counts)
    crops = [torch.from_numpy(det.det_pixels(G, 224, 224, seed + i)) for i in range(2)] + \
            [torch.from_numpy(det.det_pixels(G, 96, 96, seed + 10 + i)) for i in range(small)]
    return ([c.to(dev) for c in crops], None, [list(counts)] * (2 + small))


def test_adamw_across_the_freeze_boundary_matches_torch_adamw():
    """head.last_layer is frozen during epoch 0 (dino.py:374-376: its grads are set to None, so torch.optim.AdamW neither moves
    it nor advances its state['step']); when it is unfrozen its bias correction restarts at step 1.  Engine == autograd path +
    torch AdamW + EMA over 2 frozen + 3 unfrozen steps."""
    a, b = _build(_cfg()), _build(_cfg())
    opt = a.configure_optimizers()
    for step in range(5):
        epoch = 0 if step < 2 else 1
        for m in (a, b):
            m.current_epoch = epoch
            m.on_train_epoch_start()
        batch = _batch([1, 3, 2], 100 + step)
        opt.zero_grad(set_to_none=True)
        la = a.training_step(batch)
        la.backward()
        a.on_after_backward()
        opt.step()
        a.on_train_batch_end()
        lb = b.fused_train_step(batch)
        assert abs(la.item() - lb.item()) < 5e-3, (step, la.item(), lb.item())   # (fp32 atomics order + Adam's sign-like first steps)
    torch.cuda.synchronize()
    assert b.last_layer_steps == 3 and b.global_step == 5
    wa, wb = a.head.last_layer.weight_v.detach(), b.head.last_layer.weight_v.detach()
    moved = (wa - det_params(O.head_shapes(32, 256), 63)["last_layer.weight_v"].cuda()).abs().max().item()
    # three Adam steps of lr 1e-3 move an element by up to 3e-3; a bias correction taken at step 3..5 instead of 1..3 would
    # make the first unfrozen update ~2.7x larger (1.7e-3 off after one step)
    assert 2e-3 < moved < 3.5e-3
    d = (wa - wb).abs().flatten()               # Adam normalises updates: a few noise-level gradient elements may flip sign
    assert d.mean().item() < 5e-5 and d.float().quantile(0.99).item() < 2e-4, (d.mean().item(), d.float().quantile(0.99).item())
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        d = (p.detach() - q.detach()).abs()
        assert d.mean().item() < 1.5e-3 and d.max().item() <= 1.1e-2 * max(1.0, p.detach().abs().max().item()), k


def test_reference_lars_interface_on_the_autograd_path():
    """configure_optimizers() with optimizer.name = lars returns the reference's LARS interface (src/utils/lars.py); three
    autograd steps with it == the oracle's restatement of LARS.step applied to the same gradients."""
    opt_cfg = {"name": "lars", "lr": 0.3, "weight_decay": 1e-4, "exclude_bias_n_norm_wd": True,
               "kwargs": {"clip_lr": True, "eta": 0.02, "exclude_bias_n_norm": True, "momentum": 0.9}}
    m = _build(_cfg(opt=opt_cfg, freeze=0))
    opt = m.configure_optimizers()
    params = [p for g in opt.param_groups for p in g["params"]]
    wds = [g["weight_decay"] for g in opt.param_groups for _ in g["params"]]
    ref = [p.detach().clone().cpu() for p in params]
    bufs = [None] * len(params)
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        m.training_step(_batch([2, 1], 200 + step)).backward()
        grads = [p.grad.detach().clone().cpu() for p in params]
        opt.step()
        ref, bufs = O.lars_step(ref, grads, bufs, lr=0.3, weight_decays=wds, momentum=0.9, eta=0.02, clip_lr=True, exclude_bias_n_norm=True)
    torch.cuda.synchronize()
    worst = max(rel_err(p.detach().cpu(), r) for p, r in zip(params, ref))
    print(f"LARS autograd path: worst parameter rel err after 3 steps {worst:.2e}")
    assert worst < 5e-6
    # the bf16 shadows the kernels read were refreshed by the step
    m.backbone._ready()
    assert torch.equal(m.backbone.arena.bf16, m.backbone.arena.fp32.to(torch.bfloat16))


# Two engines fed the same batches drift apart by ~1e-4 in the loss within a few steps: gradient accumulation uses fp32 atomics
# (split-K, TMA reduce-add) whose order is not fixed, and Adam turns a noise-level gradient element into a +-lr update.  A graph
# replaying freed memory produces garbage / NaN, far outside this band.
LOSS_TOL = 3e-3


def test_cuda_graph_survives_layout_cache_eviction():
    """A captured step bakes in the device pointers of its packed layouts; evicting them from the process-wide cache (new
    ragged batches arrive all the time) must not free what the graph still reads."""
    from chadavit_b200 import ops
    g, e = _build(_cfg(graph=True)), _build(_cfg(graph=False))
    sig = [2, 1, 3]
    for step in range(3):                       # eager sighting, capture, first replay
        b = _batch(sig, 300 + step)
        lg, le = g.fused_train_step(b), e.fused_train_step(b)
        assert abs(lg.item() - le.item()) < LOSS_TOL
    assert any("graph" in ent for ent in g._graphs.values())
    rs = np.random.RandomState(0)
    for _ in range(ops.LAYOUT_CACHE_SIZE + 8):  # flood the cache: the signature's layouts are evicted
        ops.get_layout(rs.randint(1, 11, size=5).tolist(), 196, torch.device("cuda", 0))
        ops.get_layout(rs.randint(1, 11, size=5).tolist(), 36, torch.device("cuda", 0))
    assert (tuple(sig), 196, "cuda:0", 10) not in ops._LAYOUTS
    junk = [torch.full((1 << 20,), 7, dtype=torch.int32, device="cuda") for _ in range(64)]   # recycle freed blocks
    held = []
    for step in range(3, 6):
        b = _batch(sig, 300 + step)
        lg, le = g.fused_train_step(b), e.fused_train_step(b)
        held.append(lg)
        assert abs(lg.item() - le.item()) < LOSS_TOL, (step, lg.item(), le.item())
    assert len({float(h) for h in held}) == 3   # losses returned by graph replays are copies, not views of one static buffer
    del junk
    for (k, p), (_, q) in zip(g.named_parameters(), e.named_parameters()):
        d = (p.detach() - q.detach()).abs()
        # Adam turns a noise-level gradient (e.g. the key bias, whose true gradient is zero) into +-lr steps: lr = 1e-3, 6 steps
        assert d.mean().item() < 1.5e-3 and d.max().item() < 1.3e-2, k


def test_engine_state_dict_round_trip():
    """state_dict() + engine_state_dict() resume a run exactly: same losses and parameters as the uninterrupted engine."""
    a = _build(_cfg(freeze=0))
    for step in range(2):
        a.fused_train_step(_batch([1, 2], 400 + step))
    sd, esd = {k: v.clone() for k, v in a.state_dict().items()}, a.engine_state_dict()
    b = _build(_cfg(freeze=0), salt=7)
    b.load_state_dict(sd)
    b.load_engine_state_dict(esd)
    for step in range(2, 4):
        la, lb = a.fused_train_step(_batch([1, 2], 400 + step)), b.fused_train_step(_batch([1, 2], 400 + step))
        assert abs(la.item() - lb.item()) < LOSS_TOL
    assert b.global_step == 4 and abs(a.momentum_updater.cur_tau - b.momentum_updater.cur_tau) < 1e-15
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        d = (p.detach() - q.detach()).abs()
        assert d.mean().item() < 1.5e-3 and d.max().item() < 9e-3, k


@pytest.mark.parametrize("graph", [False, True])
def test_loss_to_host_handle_returns_the_device_loss(graph):
    """fused_train_step(loss_to_host=True) hands back the SAME loss through pinned memory (copied right behind the loss kernel,
    readable while the backward is still queued); eager steps and CUDA-graph replays."""
    a, b = _build(_cfg(freeze=0, graph=graph)), _build(_cfg(freeze=0, graph=graph))
    for step in range(4):            # the same batch signature: with graph=True steps 2, 3 are replays
        la = a.fused_train_step(_batch([2, 1, 3], 700 + step))
        lb = b.fused_train_step(_batch([2, 1, 3], 700 + step), loss_to_host=True)
        assert not isinstance(lb, torch.Tensor) and lb.item() == float(lb)
        if step == 0:
            assert lb.item() == la.item()                   # same parameters, deterministic forward: the very same number
        else:                                               # two engines drift in the last bits after an optimizer step (see LOSS_TOL)
            assert abs(lb.item() - la.item()) < LOSS_TOL
    # one engine: the handle and the device tensor of the SAME step (the handle's slot is filled mid-step)
    hl = b.fused_train_step(_batch([2, 1, 3], 710), loss_to_host=True)
    torch.cuda.synchronize()
    assert hl.buf.item() == hl.item()


# ---------------------------------------------------------------------------------------------------- 2 ranks, NCCL
COUNTS = [[1, 3, 2], [4, 1, 2]]     # per-rank images


def _worker(rank, world, port, out, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, ROOT)
    m = _build(_cfg(freeze=0, extra_engine={"overlap_comm": overlap, "grad_bucket_blocks": 3}), salt=rank * 100)   # rank 1 starts from OTHER weights
    losses = []
    for step in range(2):
        b = _batch(COUNTS[rank], 500 + 50 * rank + step, dev=f"cuda:{rank}")
        losses.append(m.fused_train_step(b).item())
        if step == 0:
            g_bb, g_hd, center1 = m.backbone.arena.grad.clone(), m.head.arena.grad.clone(), m.dino_loss_func.center.clone()
    torch.cuda.synchronize()
    res = {"loss": losses, "g_bb": g_bb.cpu() / world, "g_hd": g_hd.cpu() / world, "center1": center1.cpu(), "center": m.dino_loss_func.center.cpu(),
           "params": {k: v.detach().cpu() for k, v in m.named_parameters()}}
    torch.save(res, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_step_equals_single_process_on_concatenated_batch(tmp_path, overlap):
    """fused_train_step on 2 ranks (per-rank batches, bucketed gradient all-reduce + centre all-reduce on the side stream) ==
    fused_train_step of ONE process on the concatenated batch: loss mean, gradient mean, centre, parameters after two steps.
    Rank 1 is built from different weights: the start-up broadcast must make it a replica of rank 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(world, port, out, overlap), nprocs=world, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    single = _build(_cfg(freeze=0))
    losses = []
    for step in range(2):
        parts = [_batch(COUNTS[r], 500 + 50 * r + step) for r in range(world)]
        crops = [torch.cat([p[0][i] for p in parts]) for i in range(3)]
        losses.append(single.fused_train_step((crops, None, [COUNTS[0] + COUNTS[1]] * 3)).item())
        if step == 0:
            g_bb, g_hd = single.backbone.arena.grad.clone().cpu(), single.head.arena.grad.clone().cpu()
            center1 = single.dino_loss_func.center.clone().cpu()
    torch.cuda.synchronize()
    assert abs((r0["loss"][0] + r1["loss"][0]) / 2 - losses[0]) < 5e-5            # same parameters: mean of the rank losses == loss of the whole batch
    assert abs((r0["loss"][1] + r1["loss"][1]) / 2 - losses[1]) < LOSS_TOL        # after one Adam step (fp32 atomics order, see LOSS_TOL)
    assert torch.equal(r0["g_bb"], r1["g_bb"]) and torch.equal(r0["g_hd"], r1["g_hd"])          # all-reduced: bit-identical on both ranks
    assert rel_err(r0["g_bb"], g_bb) < 2e-3 and rel_err(r0["g_hd"], g_hd) < 2e-3                # == gradient of the mean loss
    # centre: after step 1 the teachers are identical, so the all-reduced batch mean must equal the single process's to rounding;
    # after step 2 the teacher has absorbed (1 - tau) of a student that took one Adam step on gradients that differ in their last
    # bits (split-K / atomic order: per-rank batches vs the concatenated one), which Adam's first step turns into +-lr (see LOSS_TOL)
    assert torch.equal(r0["center1"], r1["center1"]) and torch.equal(r0["center"], r1["center"])
    assert (r0["center1"] - center1).abs().max().item() < 1e-6
    assert (r0["center"] - single.dino_loss_func.center.cpu()).abs().max().item() < 2e-4
    for k, v in single.named_parameters():
        assert torch.equal(r0["params"][k], r1["params"][k]), k                                   # replicas stay bit-identical
        assert (r0["params"][k] - v.detach().cpu()).abs().max().item() <= 5e-3 * max(1.0, v.detach().abs().max().item()), k
