"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's ChAda-ViT / DINO hot path.

Parity status: PINNED against the reference's own modules (tests/golden/*.npz, see
oracle/__init__.py).  This file restates, with plain torch fp32 CPU arithmetic and the
reference's *padded + key-padding-mask* algorithm (NOT the packed varlen layout the CUDA
path uses), every function of SURVEY.md §8(a):

  tokenize_padded       src/backbones/vit/chada_vit.py:118-134 (TokenLearner), :219-270
  interp_pos_embed      src/backbones/vit/chada_vit.py:185-217
  encoder_layer         src/backbones/vit/chada_vit.py:75-116 (+ torch F.multi_head_attention_forward math)
  backbone_forward      src/backbones/vit/chada_vit.py:272-289
  dino_head             src/methods/dino.py:98-111 (+ weight_norm :78-81)
  dino_loss             src/losses/dino.py:69-118
  ema_update / cosine_tau   src/utils/momentum.py:63-87
  dino_step             src/methods/base.py:695-707,1216-1218 + src/methods/dino.py:279,296,313-317
  last_selfattention    src/backbones/vit/chada_vit.py:313-320 (+ :90-97 return_attention, :105-111 need_weights)
  clip_gradients        src/methods/dino.py:249-261
  lars_step             src/utils/lars.py:113-167
  one_channel_collate   src/data/channels_strategies.py:31-85
  knn_compute           src/utils/knn.py:96-177
  extract_features      src/methods/base.py:941-981 (multi_channels strategy)
  rewrite_checkpoint_keys   main_linear.py:103-110

Parameters are dicts keyed by the reference's state-dict names.  All functions are
differentiable through torch autograd, which is how the tests obtain reference gradients.
It is also the "port" CPU baseline timed by bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
FFN_DIM = 2048      # hard-coded in the reference: chada_vit.py:160
PAD_CHANNELS = 10   # forward() always pads to 10: chada_vit.py:219,274

# ---------------------------------------------------------------------------
# Optional emulation of the CUDA path's ONLY precision difference: tensor-core operands are rounded to bf16
# (activations entering a GEMM, weights, softmax probabilities); everything else (accumulation, LayerNorm, residual
# stream, softmax statistics, loss) is fp32 in both.  With OPERAND_DTYPE = None (default) this file is the plain
# fp32 restatement that is pinned to the reference.  Tests use the bf16 mode to separate "kernel is wrong" from
# "bf16 operand rounding" when comparing gradients, which are far more sensitive to rounding than the outputs.
OPERAND_DTYPE = None


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


def _q(x: Tensor) -> Tensor:
    return x if OPERAND_DTYPE is None else _RoundSTE.apply(x, OPERAND_DTYPE)


class operand_rounding:
    """Context manager: ``with operand_rounding(torch.bfloat16): ...``"""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global OPERAND_DTYPE
        self.prev, OPERAND_DTYPE = OPERAND_DTYPE, self.dtype

    def __exit__(self, *a):
        global OPERAND_DTYPE
        OPERAND_DTYPE = self.prev


# --------------------------------------------------------------------------- tokenizer
def patch_embed(x: Tensor, w: Tensor, b: Tensor, patch: int) -> Tensor:
    """TokenLearner.forward (chada_vit.py:130-134): Conv2d(1->D,k=s=patch) == unfold + GEMM.
    x (G,1,H,W) -> (G, N, D), patch index p = py*(W/patch)+px."""
    G, _, H, W = x.shape
    hp, wp = H // patch, W // patch
    cols = x[:, 0, : hp * patch, : wp * patch].reshape(G, hp, patch, wp, patch).permute(0, 1, 3, 2, 4)
    cols = cols.reshape(G, hp * wp, patch * patch)
    return _q(cols) @ _q(w.reshape(w.shape[0], -1)).t() + b


def interp_pos_embed(pos_embed: Tensor, npatch: int, w: int, h: int, patch: int) -> Tensor:
    """add_pos_encoding_per_channel(..., class_pos_embed=False) (chada_vit.py:185-217) -> (1,1,npatch,D)."""
    N = pos_embed.shape[2] - 1
    if npatch == N and w == h:
        return pos_embed[:, :, 1:]
    dim = pos_embed.shape[-1]
    w0, h0 = w // patch + 0.1, h // patch + 0.1
    s = int(math.sqrt(N))
    pp = F.interpolate(pos_embed[:, :, 1:].reshape(1, s, s, dim).permute(0, 3, 1, 2),
                       scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == pp.shape[-2] and int(h0) == pp.shape[-1]
    return pp.permute(0, 2, 3, 1).reshape(1, -1, dim).unsqueeze(0)


def tokenize_padded(x: Tensor, counts: Sequence[int], P: Dict[str, Tensor], patch: int,
                    max_channels_model: int) -> Tuple[Tensor, Tensor]:
    """channel_aware_tokenization (chada_vit.py:219-270): returns (B, 1+10N, D) and bool mask (B, 1+10N)."""
    _, _, w, h = x.shape
    tok = patch_embed(x, P["token_learner.proj.weight"], P["token_learner.proj.bias"], patch)
    N, D = tok.shape[1], tok.shape[2]
    chunks = torch.split(tok, list(counts), dim=0)
    padded = torch.stack([torch.cat([c, tok.new_zeros(PAD_CHANNELS - c.shape[0], N, D)], 0)
                          if c.shape[0] < PAD_CHANNELS else c for c in chunks], 0)      # (B,10,N,D)
    B = padded.shape[0]
    mask = (padded.reshape(B, -1, D) == 0).all(-1)                                       # :239
    padded = padded + interp_pos_embed(P["pos_embed"], N, w, h, patch)                   # :245
    if max_channels_model == PAD_CHANNELS:                                               # :248-250
        padded = padded + P["channel_token"]
    emb = padded.reshape(B, -1, D)
    cls = (P["cls_token"] + P["pos_embed"][:, :, 0]).expand(B, -1, -1)                   # :259-262
    emb = torch.cat([cls, emb], 1)
    mask = torch.cat([mask.new_zeros(B, 1), mask], 1)
    return emb, mask


# --------------------------------------------------------------------------- encoder
def mha(u: Tensor, mask: Tensor, w_in: Tensor, b_in: Tensor, w_o: Tensor, b_o: Tensor, nhead: int) -> Tensor:
    """nn.MultiheadAttention(batch_first) self-attention with a bool key-padding mask."""
    B, S, D = u.shape
    d = D // nhead
    qkv = _q(_q(u) @ _q(w_in).t() + b_in)
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(B, S, nhead, d).transpose(1, 2)
    k = k.reshape(B, S, nhead, d).transpose(1, 2)
    v = v.reshape(B, S, nhead, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    a = _q(_q(torch.softmax(s, -1)) @ v)
    return a.transpose(1, 2).reshape(B, S, D) @ _q(w_o).t() + b_o


def encoder_layer(x: Tensor, mask: Tensor, P: Dict[str, Tensor], pre: str, nhead: int, eps: float = 1e-5) -> Tensor:
    """TransformerEncoderLayer.forward, norm_first=False branch (chada_vit.py:95-100): norm1 is applied twice."""
    D = x.shape[-1]
    g1, b1 = P[pre + "norm1.weight"], P[pre + "norm1.bias"]
    u = F.layer_norm(x, (D,), g1, b1, eps)
    a = mha(u, mask, P[pre + "self_attn.in_proj_weight"], P[pre + "self_attn.in_proj_bias"],
            P[pre + "self_attn.out_proj.weight"], P[pre + "self_attn.out_proj.bias"], nhead)
    y = F.layer_norm(x + a, (D,), g1, b1, eps)
    ff = _q(torch.relu(_q(y) @ _q(P[pre + "linear1.weight"]).t() + P[pre + "linear1.bias"])) @ _q(P[pre + "linear2.weight"]).t() \
        + P[pre + "linear2.bias"]
    return F.layer_norm(y + ff, (D,), P[pre + "norm2.weight"], P[pre + "norm2.bias"], eps)


def backbone_forward(x: Tensor, index: int, list_num_channels: List[List[int]], P: Dict[str, Tensor], *,
                     nhead: int, final_eps: float, patch: int = 16, return_all_tokens: bool = False,
                     max_channels_model: int = 10, depth: int = 12) -> Tensor:
    """ChAdaViT.forward (chada_vit.py:272-289)."""
    counts = list_num_channels[index]
    h, mask = tokenize_padded(x, counts, P, patch, max_channels_model)
    for i in range(depth):
        h = encoder_layer(h, mask, P, f"blocks.{i}.", nhead)
    D = h.shape[-1]
    h = F.layer_norm(h, (D,), P["norm.weight"], P["norm.bias"], final_eps)
    if return_all_tokens:
        return h[:, 1:][~mask[:, 1:]]
    return h[:, 0]


# --------------------------------------------------------------------------- DINO head / loss / EMA
def dino_head(f: Tensor, P: Dict[str, Tensor], bn_training: bool = True) -> Tensor:
    """DINOHead.forward, num_layers=3 (src/methods/dino.py:61-111).  use_bn=False: mlp.{0,2,4} Linear; use_bn=True (detected
    from the keys): mlp.{0,3,6} Linear with BatchNorm1d mlp.{1,4} between Linear and GELU — batch statistics when
    ``bn_training`` (running statistics are updated in place in ``P`` as torch does), running statistics otherwise."""
    use_bn = "mlp.6.weight" in P
    lin = ("mlp.0", "mlp.3", "mlp.6") if use_bn else ("mlp.0", "mlp.2", "mlp.4")

    def bn(h, pre):
        if not use_bn:
            return h
        rm, rv = P[pre + ".running_mean"], P[pre + ".running_var"]
        if bn_training:
            mean, var = h.mean(0), h.var(0, unbiased=False)
            with torch.no_grad():
                n = h.shape[0]
                rm.mul_(0.9).add_(0.1 * mean.detach())
                rv.mul_(0.9).add_(0.1 * var.detach() * n / max(n - 1, 1))
                if pre + ".num_batches_tracked" in P:
                    P[pre + ".num_batches_tracked"] += 1
        else:
            mean, var = rm, rv
        return (h - mean) / torch.sqrt(var + 1e-5) * P[pre + ".weight"] + P[pre + ".bias"]
    h = F.gelu(bn(f @ P[lin[0] + ".weight"].t() + P[lin[0] + ".bias"], "mlp.1"))
    h = F.gelu(bn(h @ P[lin[1] + ".weight"].t() + P[lin[1] + ".bias"], "mlp.4"))
    h = h @ P[lin[2] + ".weight"].t() + P[lin[2] + ".bias"]
    h = h / h.norm(dim=-1, keepdim=True).clamp_min(1e-12)                                # F.normalize
    v, g = P["last_layer.weight_v"], P["last_layer.weight_g"]
    w = v * (g / v.norm(dim=1, keepdim=True))                                            # weight_norm, dim=0
    return h @ w.t()


def teacher_temp(epoch: int, warmup_teacher_temp: float, teacher_temp_: float, warmup_epochs: int) -> float:
    """teacher_temp_schedule[epoch] (src/losses/dino.py:62-67): linspace warm-up then constant."""
    if epoch < warmup_epochs:
        if warmup_epochs == 1:
            return float(warmup_teacher_temp)
        return float(warmup_teacher_temp + (teacher_temp_ - warmup_teacher_temp) * epoch / (warmup_epochs - 1))
    return float(teacher_temp_)


def dino_loss(student: Tensor, teacher: Tensor, center: Tensor, *, student_temp: float, teacher_temp: float,
              num_large_crops: int = 2, center_momentum: float = 0.9, world_size: int = 1,
              all_reduce_sum=None) -> Tuple[Tensor, Tensor]:
    """DINOLoss.forward + update_center (src/losses/dino.py:69-118). Returns (loss, new_center)."""
    s = (student / student_temp).chunk(num_large_crops)
    q = torch.softmax((teacher - center) / teacher_temp, -1).detach().chunk(2)
    total, n = 0.0, 0
    for iq, qq in enumerate(q):
        for iv, v in enumerate(s):
            if iv == iq:
                continue
            total = total + torch.sum(-qq * F.log_softmax(v, -1), -1).mean()
            n += 1
    total = total / n
    with torch.no_grad():
        bc = teacher.sum(0, keepdim=True)
        if all_reduce_sum is not None:
            bc = all_reduce_sum(bc)
        bc = bc / world_size / len(teacher)
        new_center = center * center_momentum + bc * (1 - center_momentum)
    return total, new_center


def ema_update(online: Sequence[Tensor], momentum: Sequence[Tensor], tau: float) -> List[Tensor]:
    """MomentumUpdater.update (src/utils/momentum.py:73-74)."""
    return [tau * m + (1 - tau) * o for o, m in zip(online, momentum)]


def cosine_tau(base_tau: float, final_tau: float, cur_step: int, max_steps: int) -> float:
    """MomentumUpdater.update_tau (src/utils/momentum.py:84-87)."""
    return final_tau - (final_tau - base_tau) * (math.cos(math.pi * cur_step / max_steps) + 1) / 2


def dino_step(crops: Sequence[Tensor], list_num_channels: List[List[int]], student: Dict[str, Tensor],
              student_head: Dict[str, Tensor], teacher: Dict[str, Tensor], teacher_head: Dict[str, Tensor],
              center: Tensor, *, nhead: int, final_eps: float, num_large_crops: int = 2, student_temp: float = 0.1,
              teacher_temp: float = 0.07, run_local_crops: bool = True) -> Tuple[Tensor, Tensor]:
    """One DINO training-step forward with the reference wiring: the student head and the
    loss see only the large crops; small crops go through the student backbone and are discarded.
    Returns (loss, new_center)."""
    z = []
    for i in range(num_large_crops):
        f = backbone_forward(crops[i], i, list_num_channels, student, nhead=nhead, final_eps=final_eps)
        z.append(dino_head(f, student_head))
    if run_local_crops:
        for i, xc in enumerate(crops[num_large_crops:]):   # base.py:701-707: index restarts at 0 (Q12)
            backbone_forward(xc, i, list_num_channels, student, nhead=nhead, final_eps=final_eps)
    with torch.no_grad():
        mz = []
        for i in range(num_large_crops):
            f = backbone_forward(crops[i], i, list_num_channels, teacher, nhead=nhead, final_eps=final_eps)
            mz.append(dino_head(f, teacher_head))
    return dino_loss(torch.cat(z), torch.cat(mz), center, student_temp=student_temp, teacher_temp=teacher_temp,
                     num_large_crops=num_large_crops)


# --------------------------------------------------------------------------- helpers for tests
def backbone_shapes(D: int, depth: int = 12, patch: int = 16, npatch: int = 196, max_ch: int = 10) -> Dict[str, tuple]:
    """State-dict names/shapes of ChAdaViT, in the reference's registration order."""
    s: Dict[str, tuple] = {
        "cls_token": (1, 1, D), "channel_token": (1, max_ch, 1, D), "pos_embed": (1, 1, npatch + 1, D),
        "token_learner.proj.weight": (D, 1, patch, patch), "token_learner.proj.bias": (D,),
    }
    for i in range(depth):
        p = f"blocks.{i}."
        s[p + "self_attn.in_proj_weight"] = (3 * D, D)
        s[p + "self_attn.in_proj_bias"] = (3 * D,)
        s[p + "self_attn.out_proj.weight"] = (D, D)
        s[p + "self_attn.out_proj.bias"] = (D,)
        s[p + "linear1.weight"] = (FFN_DIM, D)
        s[p + "linear1.bias"] = (FFN_DIM,)
        s[p + "linear2.weight"] = (D, FFN_DIM)
        s[p + "linear2.bias"] = (D,)
        s[p + "norm1.weight"] = (D,)
        s[p + "norm1.bias"] = (D,)
        s[p + "norm2.weight"] = (D,)
        s[p + "norm2.bias"] = (D,)
    s["norm.weight"] = (D,)
    s["norm.bias"] = (D,)
    return s


def head_shapes(in_dim: int, K: int, hidden: int = 2048, bottleneck: int = 256, use_bn: bool = False) -> Dict[str, tuple]:
    """state_dict shapes of DINOHead(in_dim, K, use_bn=use_bn) in registration order (buffers of BatchNorm1d included)."""
    if not use_bn:
        return {"mlp.0.weight": (hidden, in_dim), "mlp.0.bias": (hidden,), "mlp.2.weight": (hidden, hidden), "mlp.2.bias": (hidden,),
                "mlp.4.weight": (bottleneck, hidden), "mlp.4.bias": (bottleneck,),
                "last_layer.weight_g": (K, 1), "last_layer.weight_v": (K, bottleneck)}
    d: Dict[str, tuple] = {}
    for lin, bn, (o, i) in (("mlp.0", "mlp.1", (hidden, in_dim)), ("mlp.3", "mlp.4", (hidden, hidden))):
        d[lin + ".weight"], d[lin + ".bias"] = (o, i), (o,)
        d[bn + ".weight"], d[bn + ".bias"], d[bn + ".running_mean"], d[bn + ".running_var"] = (o,), (o,), (o,), (o,)
        d[bn + ".num_batches_tracked"] = ()
    d["mlp.6.weight"], d["mlp.6.bias"] = (bottleneck, hidden), (bottleneck,)
    d["last_layer.weight_g"], d["last_layer.weight_v"] = (K, 1), (K, bottleneck)
    return d


def packed_index(counts: Sequence[int], npatch: int) -> Tuple[List[int], List[int]]:
    """cu_seqlens and, for every packed row, its row in the reference's padded (B,1+10N) layout."""
    cu, rows = [0], []
    S_pad = 1 + PAD_CHANNELS * npatch
    for b, c in enumerate(counts):
        n = 1 + c * npatch
        rows.extend(b * S_pad + r for r in range(n))
        cu.append(cu[-1] + n)
    return cu, rows


# --------------------------------------------------------------------------- §8(f) rows: attention maps, clip, LARS, collate
def last_selfattention(x: Tensor, P: Dict[str, Tensor], *, nhead: int, patch: int = 16, depth: int = 12) -> Tensor:
    """ChAdaViT.get_last_selfattention (chada_vit.py:313-320): every channel image is its own single-channel sequence
    (``list_num_channels=[1]`` makes ``torch.split`` cut chunks of one; ``max_channels=1`` means no padding and — because
    1 != self.max_channels — no channel token); blocks 0..depth-2 run normally; the last block returns the attention
    probabilities of ``self_attn(norm1(x))`` per head (``average_attn_weights=False``): (ΣC, nhead, 1+N, 1+N)."""
    _, _, w, h = x.shape
    tok = patch_embed(x, P["token_learner.proj.weight"], P["token_learner.proj.bias"], patch)     # (G, N, D)
    G, N, D = tok.shape
    mask = torch.cat([tok.new_zeros(G, 1, dtype=torch.bool), (tok == 0).all(-1)], 1)              # :239,267 (all False in practice)
    emb = tok + interp_pos_embed(P["pos_embed"], N, w, h, patch)[:, 0]
    cls = (P["cls_token"] + P["pos_embed"][:, :, 0]).expand(G, -1, -1)
    hcur = torch.cat([cls, emb], 1)
    for i in range(depth - 1):
        hcur = encoder_layer(hcur, mask, P, f"blocks.{i}.", nhead)
    pre = f"blocks.{depth - 1}."
    u = F.layer_norm(hcur, (D,), P[pre + "norm1.weight"], P[pre + "norm1.bias"], 1e-5)
    S, d = u.shape[1], D // nhead
    qkv = _q(_q(u) @ _q(P[pre + "self_attn.in_proj_weight"]).t() + P[pre + "self_attn.in_proj_bias"])
    q, k, _v = qkv.split(D, dim=-1)
    q = q.reshape(G, S, nhead, d).transpose(1, 2)
    k = k.reshape(G, S, nhead, d).transpose(1, 2)
    s_ = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    s_ = s_.masked_fill(mask[:, None, None, :], float("-inf"))
    return torch.softmax(s_, -1)


def clip_gradients(grads: Sequence[Tensor], clip: float) -> List[Tensor]:
    """DINO.dino_clip_gradients (dino.py:249-261): per-parameter clip_coef = clip / (||g||_2 + 1e-6), applied when < 1."""
    out = []
    for g in grads:
        if g is None:
            out.append(None)
            continue
        coef = clip / (g.norm(2) + 1e-6)
        out.append(g * coef if coef < 1 else g.clone())
    return out


def lars_step(params: Sequence[Tensor], grads: Sequence[Tensor], bufs: List, *, lr: float, weight_decays: Sequence[float],
              momentum: float = 0.0, dampening: float = 0.0, nesterov: bool = False, eta: float = 1e-3, eps: float = 1e-8,
              clip_lr: bool = False, exclude_bias_n_norm: bool = False) -> Tuple[List[Tensor], List]:
    """LARS.step (src/utils/lars.py:113-167) for a flat list of parameters; ``weight_decays[i]`` is the weight decay of the
    group parameter i sits in (base.py:426-427 puts ndim <= 1 parameters into a weight_decay = 0 group when
    exclude_bias_n_norm_wd).  ``grads[i] is None`` skips the parameter (:128-129).  Returns (new params, new buffers)."""
    new_p, new_b = [], []
    for p, g, buf, wd in zip(params, grads, bufs, weight_decays):
        if g is None:
            new_p.append(p.clone())
            new_b.append(buf)
            continue
        d_p = g
        p_norm, g_norm = torch.norm(p), torch.norm(g)
        if p.ndim != 1 or not exclude_bias_n_norm:                       # :136
            if p_norm != 0 and g_norm != 0:
                lars_lr = p_norm / (g_norm + p_norm * wd + eps)          # :138
                lars_lr = lars_lr * eta
                if clip_lr:
                    lars_lr = min(lars_lr / lr, 1)                       # :143
                d_p = d_p.add(p, alpha=wd)
                d_p = d_p * lars_lr
        if momentum != 0:                                                # :149-159
            buf = d_p.clone() if buf is None else buf * momentum + (1 - dampening) * d_p
            d_p = d_p.add(buf, alpha=momentum) if nesterov else buf
        new_p.append(p - lr * d_p)
        new_b.append(buf)
    return new_p, new_b


def one_channel_collate(batch):
    """one_channel_collate_fn (channels_strategies.py:31-85): (crop_lists, labels, num_channels_lists)."""
    first = batch[0][-2:][0]
    num_crops = len(first) if isinstance(first, list) else 1
    crop_lists = [[] for _ in range(num_crops)]
    counts = [[] for _ in range(num_crops)]
    labels = []
    for item in batch:
        image_list, label = item[-2:]
        if isinstance(image_list, torch.Tensor):
            image_list = [image_list]
        for i, crop in enumerate(image_list):
            counts[i].append(crop.shape[0])
            for c in range(crop.shape[0]):
                crop_lists[i].append(crop[c, :, :].unsqueeze(0))
        labels.append(label)
    crop_lists = [torch.cat(c, 0).unsqueeze(1) for c in crop_lists]
    return (crop_lists[0] if len(crop_lists) == 1 else crop_lists), torch.tensor(labels), counts


def knn_compute(train_features: Tensor, train_targets: Tensor, test_features: Tensor, test_targets: Tensor, *, k: int = 20,
                T: float = 0.07, max_distance_matrix_size: int = int(5e6), distance_fx: str = "cosine",
                epsilon: float = 0.00001) -> Tuple[float, float, Tensor]:
    """WeightedKNNClassifier.compute (knn.py:96-177) -> (top1 %, top5 %, predicted class per test sample)."""
    if distance_fx == "cosine":
        train_features = F.normalize(train_features)
        test_features = F.normalize(test_features)
    num_classes = torch.unique(test_targets).numel()
    n_train, n_test = train_targets.size(0), test_targets.size(0)
    chunk = min(max(1, max_distance_matrix_size // n_train), n_test)
    k = min(k, n_train)
    top1 = top5 = total = 0.0
    preds = []
    for idx in range(0, n_test, chunk):
        feats = test_features[idx:min(idx + chunk, n_test)]
        targets = test_targets[idx:min(idx + chunk, n_test)]
        b = targets.size(0)
        if distance_fx == "cosine":
            sim = feats @ train_features.t()
        elif distance_fx == "euclidean":
            sim = 1 / (torch.cdist(feats, train_features) + epsilon)
        else:
            raise NotImplementedError
        sim, ind = sim.topk(k, largest=True, sorted=True)
        neigh = torch.gather(train_targets.view(1, -1).expand(b, -1), 1, ind)
        one_hot = torch.zeros(b * k, num_classes)
        one_hot.scatter_(1, neigh.view(-1, 1).long(), 1)
        if distance_fx == "cosine":
            sim = (sim / T).exp()
        probs = (one_hot.view(b, -1, num_classes) * sim.view(b, -1, 1)).sum(1)
        _, pred = probs.sort(1, True)
        correct = pred.eq(targets.view(-1, 1))
        top1 += correct[:, :1].sum().item()
        top5 += correct[:, :min(5, k, correct.size(-1))].sum().item()
        total += b
        preds.append(pred[:, 0])
    return top1 * 100.0 / total, top5 * 100.0 / total, torch.cat(preds)


def extract_features(feats: Tensor, counts: Sequence[int], *, return_all_tokens: bool, mixed_channels: bool = False) -> Tensor:
    """Tail of BaseMethod._base_extract_step for the multi_channels strategy (base.py:958-981): with return_all_tokens the
    (ΣC·N, D) token matrix becomes one row per image (needs equal channel counts: torch.stack); CLS features pass through."""
    if mixed_channels or not return_all_tokens:
        return feats
    chunks = feats.view(sum(counts), -1, feats.shape[-1])
    chunks = torch.split(chunks, list(counts), dim=0)
    return torch.stack(chunks, dim=0).flatten(start_dim=1)


def rewrite_checkpoint_keys(state: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Key rewriting of main_linear.py:103-110 applied to a Lightning ``state_dict``: 'encoder' -> 'backbone', then the
    'backbone.' prefix is stripped; every original key is deleted (so only backbone entries survive)."""
    state = dict(state)
    for k in list(state.keys()):
        if "encoder" in k:
            state[k.replace("encoder", "backbone")] = state[k]
        if "backbone" in k:
            state[k.replace("backbone.", "")] = state[k]
        del state[k]
    return state
