"""CPU: host-side bookkeeping (bit-exact index contract), module surface / state-dict parity, schedules, and the
world_size-2 (gloo) equivalence of the data-parallel reductions."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import chada_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cls_tail_applies_only_to_cls_backbones():
    """The CLS-only last block is taken only when the backbone returns x[:, 0] (chada_vit.py:289) and there is a block in front of it."""
    from chadavit_b200 import ops
    assert ops.cls_tail_ok(12, False, 96) and ops.cls_tail_ok(2, False, 16)
    assert not ops.cls_tail_ok(12, True, 96)        # return_all_tokens: every row of the last block is an output
    assert not ops.cls_tail_ok(1, False, 96)


@pytest.mark.parametrize("counts,npatch", [([1, 3, 5, 10], 196), ([2, 10, 1], 36), ([10] * 7, 196), ([1], 4)])
def test_packed_layout_is_bit_exact(counts, npatch):
    """cu_seqlens / channel->image maps / row order == the reference's split-pad-stack order at the unmasked positions."""
    from chadavit_b200.ops import PackedLayout
    lay = PackedLayout(counts, npatch, "cpu")
    cu, rows = O.packed_index(counts, npatch)
    assert lay.cu_host.tolist() == cu and lay.T == len(rows) == len(counts) + sum(counts) * npatch
    # row cu[b]+1+c*N+p  <->  channel image (off_b + c), patch p
    off = np.concatenate([[0], np.cumsum(counts)])
    for b, C in enumerate(counts):
        for c in (0, C - 1):
            g = off[b] + c
            assert lay.chan_img_host[g] == b and lay.chan_idx_host[g] == c
            assert g * npatch + 0 + b + 1 == cu[b] + 1 + c * npatch      # packed row of patch 0 of channel image g
    # reference mask: padded rows are exactly the rows NOT in `rows`
    S_pad = 1 + 10 * npatch
    masked = np.ones(len(counts) * S_pad, bool)
    masked[rows] = False
    for b, C in enumerate(counts):
        assert masked[b * S_pad:(b + 1) * S_pad].sum() == (10 - C) * npatch   # channel_mask.sum(1) == 1960 - 196*C
    w = lay.attn_work(2).numpy()
    assert w.shape[1] == 4 and len(w) == 2 * sum((cu[b + 1] - cu[b] + 127) // 128 for b in range(len(counts)))
    lens = w[:, 2] - w[:, 1]
    assert (np.diff(lens) <= 0).all()                                 # longest sequences first
    nc = lay.non_cls_rows().numpy()
    assert len(nc) == sum(counts) * npatch and not set(nc) & set(cu[:-1])
    c64 = lay.cls_rows64()                                            # CLS rows for the CLS-only last block (chada_vit.py:289)
    assert c64.dtype == torch.int64 and c64.tolist() == cu[:-1]
    with pytest.raises(ValueError):
        PackedLayout([11], npatch, "cpu")
    with pytest.raises(ValueError):
        PackedLayout([0, 2], npatch, "cpu")


def test_module_surface_matches_reference_contract():
    from chadavit_b200.backbones import ChAdaViT, chada_vit, vit_channels
    from chadavit_b200.methods import DINOHead
    m = vit_channels("dino", patch_size=16, embed_dim=192, return_all_tokens=False, max_number_channels=10, ignored="x")
    assert isinstance(m, ChAdaViT) and m.num_heads == 2 and m.norm.eps == 1e-6 and len(m.blocks) == 12
    assert m.num_features == m.embed_dim == 192 and m.token_learner.num_patches == 196 and m.token_learner.patch_size == 16
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == O.backbone_shapes(192)
    assert list(sd.keys()) == list(O.backbone_shapes(192).keys())
    assert sum(p.numel() for p in m.parameters()) == 11_341_632
    bare = ChAdaViT(patch_size=16, embed_dim=192, return_all_tokens=False, max_number_channels=10)
    assert bare.num_heads == 12 and bare.norm.eps == 1e-5                # Q4 / Q5
    bare.load_state_dict(sd)                                             # same weights load either way (Q4)
    h = DINOHead(192, 4096, use_bn=False)
    assert {k: tuple(v.shape) for k, v in h.state_dict().items()} == O.head_shapes(192, 4096)
    assert sum(p.numel() for p in h.parameters()) == 6_168_832 and not h.last_layer.weight_g.requires_grad
    hb = DINOHead(192, 4096)                                             # the class default use_bn=True (dino.py:40,66-73)
    assert {k: tuple(v.shape) for k, v in hb.state_dict().items()} == O.head_shapes(192, 4096, use_bn=True)
    # flat arena keeps parameters as views, survives load_state_dict and is rebuilt after .data is replaced
    a = m.arena
    a.ensure()
    p = m.blocks[3].linear1.weight
    assert p.data_ptr() == a.fp32.data_ptr() + 4 * a.offsets["blocks.3.linear1.weight"][0]
    p.data = p.data.clone()
    a.ensure()
    assert p.data_ptr() == a.fp32.data_ptr() + 4 * a.offsets["blocks.3.linear1.weight"][0]


def test_schedules_match_reference_formulas():
    from chadavit_b200.losses import DINOLoss
    from chadavit_b200.utils.momentum import MomentumUpdater
    L = DINOLoss(64, 0.04, 0.07, 3, 10)
    assert np.allclose(L.teacher_temp_schedule[:4], [0.04, 0.055, 0.07, 0.07]) and len(L.teacher_temp_schedule) == 10
    for e in range(10):
        assert abs(L.teacher_temp_schedule[e] - O.teacher_temp(e, 0.04, 0.07, 3)) < 1e-12
    u = MomentumUpdater(0.9995, 1.0)
    for step in (0, 10, 50, 100):
        u.update_tau(step, 100)
        assert abs(u.cur_tau - O.cosine_tau(0.9995, 1.0, step, 100)) < 1e-15
    with pytest.raises(AssertionError):
        MomentumUpdater(1.0, 0.5)


def test_interp_matrix_equals_reference_bicubic():
    from chadavit_b200.backbones import chada_vit
    m = chada_vit(patch_size=16, embed_dim=32, return_all_tokens=False, max_number_channels=10)
    M = m._interp_matrix(6, 6, 96, 96, "cpu")
    ref = O.interp_pos_embed(m.pos_embed.detach(), 36, 96, 96, 16)[0, 0]
    assert (M @ m.pos_embed.detach()[0, 0, 1:] - ref).abs().max().item() < 1e-6


# ---------------------------------------------------------------------------------------------------- world_size 2 (gloo)
def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import det
    B, K, V = 3, 64, 2
    s = torch.from_numpy(det.det_uniform((world, V * B, K), 5))[rank].clone().requires_grad_()
    t = torch.from_numpy(det.det_uniform((world, 2 * B, K), 6))[rank]

    def allred(x):
        dist.all_reduce(x)
        return x
    loss, center = O.dino_loss(s, t, torch.zeros(1, K), student_temp=0.1, teacher_temp=0.07, world_size=world, all_reduce_sum=allred)
    loss.backward()
    g = s.grad.clone()
    dist.all_reduce(g)                       # DDP: gradient of the mean-over-ranks loss wrt a replicated parameter ~ averaged grads
    lm = loss.detach().clone()
    dist.all_reduce(lm)
    if rank == 0:
        torch.save({"center": center, "loss_mean": lm / world}, out)
    dist.destroy_process_group()


def test_two_rank_reductions_equal_single_process_on_concatenated_batch(tmp_path):
    """C2 (centre all-reduce: sum -> /world -> /rows) and the loss mean over ranks reproduce a single-process run on the
    concatenated batch (SURVEY.md §8e equivalence test), exercised with gloo on CPU."""
    from oracle import det
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    B, K, V = 3, 64, 2
    s = torch.from_numpy(det.det_uniform((world, V * B, K), 5))
    t = torch.from_numpy(det.det_uniform((world, 2 * B, K), 6))
    # concatenated batch: view v of the big batch = cat over ranks of view v
    S = torch.cat([torch.cat([s[r, v * B:(v + 1) * B] for r in range(world)]) for v in range(V)])
    T = torch.cat([torch.cat([t[r, v * B:(v + 1) * B] for r in range(world)]) for v in range(2)])
    loss, center = O.dino_loss(S, T, torch.zeros(1, K), student_temp=0.1, teacher_temp=0.07)
    assert (got["center"] - center).abs().max().item() < 1e-7
    assert abs(got["loss_mean"].item() - loss.item()) < 1e-6


def test_bench_sharding_is_token_balanced():
    """bench.py deals every step's GLOBAL batch (same draw on every rank) to the ranks with data/balance.py: the shards are a
    partition of the global batch with equal cardinality, and the most loaded rank carries < 1 % more estimated work than the
    mean (contiguous DistributedSampler-like slices: several %)."""
    sys.path.insert(0, ROOT)
    import bench
    from chadavit_b200.data.balance import image_cost, token_balanced_shards
    c0, c1 = bench.channel_counts(64), bench.channel_counts(64)
    assert c0 == c1 and len(c0) == 64 and min(c0) >= 1 and max(c0) <= 10
    for world in (2, 4, 8):
        for step in range(3):
            glob = bench.channel_counts(64 * world, seed=1234 + 7919 * step)
            shards = token_balanced_shards(glob, world)
            assert sorted(i for s in shards for i in s) == list(range(64 * world)) and all(len(s) == 64 for s in shards)
    worst_bal, worst_raw = 0.0, 0.0
    for step in range(6):                       # bench.py: rank r runs shard r of the 8-GPU job's global batch, at every N
        glob = bench.channel_counts(64 * bench.SHARD_WORLD, seed=1234 + 7919 * step)
        per_rank = [bench.step_counts(step, r, 8) for r in range(8)]
        assert sorted(c for cs in per_rank for c in cs) == sorted(glob)
        assert bench.step_counts(step, 1, 2) == per_rank[1] and bench.step_counts(step, 0, 1) == per_rank[0]
        load = np.array([sum(image_cost(c) for c in cs) for cs in per_rank])
        raw = np.array([sum(image_cost(c) for c in bench.step_counts(step, r, 8, balanced=False)) for r in range(8)])
        worst_bal, worst_raw = max(worst_bal, load.max() / load.mean()), max(worst_raw, raw.max() / raw.mean())
    assert worst_bal < 1.01 < worst_raw, (worst_bal, worst_raw)
    with pytest.raises(ValueError):
        token_balanced_shards([1, 2, 3], 2)


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from chadavit_b200.methods import DINO
    m = DINO({"backbone": {"kwargs": {"embed_dim": 32}}, "method_kwargs": {"num_prototypes": 64}, "engine": {"grad_bucket_blocks": 5}})
    a = m.backbone.arena
    g = torch.arange(a.numel, dtype=torch.float32) * (rank + 1)
    flat = g.clone()
    dist.all_reduce(flat)
    fired = []
    for i in list(range(m.backbone.depth - 1, -1, -1)) + [-1]:        # the order _backward_impl calls block_done in
        for lo, hi, at in m._grad_buckets():
            if at == i:
                dist.all_reduce(g[lo:hi])
                fired.append((lo, hi, at))
    if rank == 0:
        torch.save({"equal": torch.equal(g, flat), "fired": fired, "numel": a.numel}, out)
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_plan_two_ranks(tmp_path):
    """The engine's gradient buckets (DINO._grad_buckets: contiguous arena slices of `grad_bucket_blocks` encoder blocks, fired
    as the backward retires them) partition the arena, fire in descending address order, contain every parameter of the blocks
    differentiated so far, and all-reducing them one by one (gloo, world_size 2) equals ONE flat all-reduce."""
    from chadavit_b200.methods import DINO
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "b.pt")
    mp.spawn(_bucket_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    assert got["equal"]
    fired = got["fired"]
    assert fired[0][1] == got["numel"] and fired[-1][0] == 0 and fired[-1][2] == -1
    assert all(fired[k][0] == fired[k + 1][1] for k in range(len(fired) - 1))            # contiguous, descending, no gaps
    for nb in (1, 3, 12, 40):
        m = DINO({"backbone": {"kwargs": {"embed_dim": 32}}, "method_kwargs": {"num_prototypes": 64}, "engine": {"grad_bucket_blocks": nb}})
        a = m.backbone.arena
        bk = m._grad_buckets()
        assert sum(hi - lo for lo, hi, _ in bk) == a.numel and 1 <= len(bk) <= 13
        for lo, hi, at in bk:
            for n in a.names:
                off, cnt, _ = a.offsets[n]
                if lo <= off < hi:                                   # a parameter inside a bucket is complete when the bucket fires
                    blk = int(n.split(".")[1]) if n.startswith("blocks.") else (99 if n.startswith("norm.") else -1)
                    assert blk >= at, (n, at)
                    assert off + cnt <= hi


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("kind,tile", [("fwd", 256), ("fwd", 128), ("bwd", 128)])
def test_attn_schedule_is_a_permutation_and_balanced(seed, kind, tile):
    """The LPT schedule the persistent attention kernels walk (CTA c takes slots c, c + G, ...): every work item of the
    plain list appears exactly once, padding slots are empty (seq_end <= seq_start), and the most loaded CTA carries at most
    4/3 of the mean load + one item (Graham's bound for LPT) — pure host arithmetic, random ragged batches."""
    from chadavit_b200.ops import PackedLayout
    rs = np.random.RandomState(seed)
    counts = rs.randint(1, 11, size=int(rs.randint(3, 70))).tolist()
    npatch = int(rs.choice([36, 196]))
    lay = PackedLayout(counts, npatch, "cpu")
    heads = int(rs.choice([2, 12]))
    for G in (8, 148):
        work = lay.attn_work(heads, tile).numpy()
        sched = lay.attn_schedule(heads, tile, kind, n_ctas=G).numpy()
        real = sched[sched[:, 2] > sched[:, 1]]
        assert sorted(map(tuple, real)) == sorted(map(tuple, work))
        assert (sched[sched[:, 2] <= sched[:, 1]] == 0).all()
        if len(work) <= G:
            continue
        assert len(sched) % G == 0
        seq = (real[:, 2] - real[:, 1]).astype(np.int64)
        if kind == "fwd":
            cost = ((seq + 63) // 64) * np.minimum((real[:, 2] - real[:, 0] + 127) // 128, tile // 128) + 3
        else:
            cost = (seq + 127) // 128 + 1
        loads = np.zeros(G, dtype=np.int64)
        slot_cta = np.nonzero(sched[:, 2] > sched[:, 1])[0] % G
        np.add.at(loads, slot_cta, cost)
        assert loads.max() <= 4 / 3 * loads.mean() + cost.max()


def test_overlay_tree_has_the_reference_module_paths():
    """The ``src/`` overlay ("ships its own src/backbones/vit/chada_vit.py ... same module path"): every name the
    reference imports from these modules resolves to the chadavit_b200 class, the factory builds an instance of THE class
    ``isinstance`` is checked against (base.py:526), and — when the reference checkout is at hand — constructor / method
    signatures equal the reference's (parsed, never imported)."""
    import ast
    import importlib
    import inspect
    sys.path.insert(0, ROOT)
    want = {"src.backbones.vit.chada_vit": ["ChAdaViT", "TransformerEncoderLayer", "TokenLearner", "chada_vit"],
            "src.losses.dino": ["DINOLoss"], "src.utils.momentum": ["MomentumUpdater", "initialize_momentum_params"],
            "src.utils.lars": ["LARS"], "src.methods.dino": ["DINOHead"]}
    impl = {"src.backbones.vit.chada_vit": "chadavit_b200.backbones.chada_vit", "src.losses.dino": "chadavit_b200.losses.dino",
            "src.utils.momentum": "chadavit_b200.utils.momentum", "src.utils.lars": "chadavit_b200.utils.lars",
            "src.methods.dino": "chadavit_b200.methods.dino"}
    for mod, names in want.items():
        m, real = importlib.import_module(mod), importlib.import_module(impl[mod])
        for n in names:
            assert getattr(m, n) is getattr(real, n), (mod, n)
    from src.backbones import vit_channels
    from src.backbones.vit.chada_vit import ChAdaViT
    bb = vit_channels("dino", patch_size=16, embed_dim=32, return_all_tokens=False, max_number_channels=10)
    assert isinstance(bb, ChAdaViT) and bb.num_heads == 2 and bb.norm.eps == 1e-6          # chada_vit.py:333-339
    ref_root = "/root/reference/src"
    if not os.path.isdir(ref_root):
        pytest.skip("reference checkout not present: signature comparison skipped")

    def ref_sig(path, cls, fn):
        tree = ast.parse(open(os.path.join(ref_root, path)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ClassDef) and node.name == cls:
                for f in node.body:
                    if isinstance(f, ast.FunctionDef) and f.name == fn:
                        return [a.arg for a in f.args.args]
        raise AssertionError((path, cls, fn))
    for path, mod, cls, fn in [("backbones/vit/chada_vit.py", "src.backbones.vit.chada_vit", "ChAdaViT", "__init__"),
                               ("backbones/vit/chada_vit.py", "src.backbones.vit.chada_vit", "ChAdaViT", "forward"),
                               ("losses/dino.py", "src.losses.dino", "DINOLoss", "__init__"),
                               ("losses/dino.py", "src.losses.dino", "DINOLoss", "forward"),
                               ("utils/momentum.py", "src.utils.momentum", "MomentumUpdater", "__init__"),
                               ("utils/momentum.py", "src.utils.momentum", "MomentumUpdater", "update"),
                               ("utils/momentum.py", "src.utils.momentum", "MomentumUpdater", "update_tau"),
                               ("utils/lars.py", "src.utils.lars", "LARS", "__init__"),
                               ("methods/dino.py", "src.methods.dino", "DINOHead", "__init__"),
                               ("methods/dino.py", "src.methods.dino", "DINOHead", "forward")]:
        ours = list(inspect.signature(getattr(getattr(importlib.import_module(mod), cls), fn)).parameters)
        ours = [a for a in ours if a != "kwargs"]
        assert ours == ref_sig(path, cls, fn), (cls, fn, ours, ref_sig(path, cls, fn))
