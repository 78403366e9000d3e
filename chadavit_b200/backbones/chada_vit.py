"""B200-native ChAda-ViT backbone behind the reference's API (src/backbones/vit/chada_vit.py).

Same class names, constructor, ``forward(x, index, list_num_channels)`` contract, attributes and ``state_dict`` keys as
the reference (chada_vit.py:136-339), but the arithmetic runs on a PACKED varlen token buffer through the sm_100a
kernels of libchadavit_b200 (tokenizer GEMM with fused embedding epilogue, tcgen05 GEMMs, varlen flash-style
attention, LayerNorm, and their backward kernels).  The torch submodules (``nn.Conv2d``, ``nn.MultiheadAttention``,
``nn.Linear``, ``nn.LayerNorm``) are kept ONLY as parameter containers so that names, shapes, registration order and the
default initialisation are the reference's; their ``forward`` is never called.  There is no CPU path.

Block semantics reproduced exactly: ``a = MHA(norm1(x)); x = norm1(x + a); x = norm2(x + W2 relu(W1 x))``.
"""
from __future__ import annotations

import math
import os
from functools import partial
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..arena import ParamArena

FFN_DIM = 2048      # chada_vit.py:160 (dim_feedforward hard-coded)
_MASK_BITS = os.environ.get("CB_NO_MASK_BITS", "") != "1"   # A/B switch: d(hidden) masked by the 1-bit mask (default) or by the bf16 activations
PAD_CHANNELS = 10   # chada_vit.py:219 (forward always pads to 10; here: the upper bound on channels per image)


def trunc_normal_(tensor: torch.Tensor, mean: float = 0.0, std: float = 1.0, a: float = -2.0, b: float = 2.0) -> torch.Tensor:
    """Same algorithm as src/utils/misc.py:134-178 (inverse-CDF truncated normal)."""
    def norm_cdf(v):
        return (1.0 + math.erf(v / math.sqrt(2.0))) / 2.0
    with torch.no_grad():
        lo, hi = norm_cdf((a - mean) / std), norm_cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * hi - 1).erfinv_().mul_(std * math.sqrt(2.0)).add_(mean).clamp_(min=a, max=b)
    return tensor


class TransformerEncoderLayer(nn.Module):
    """Parameter container with the reference layer's attribute names (chada_vit.py:29-116).  The computation
    lives in ChAdaViT._block_fwd/_block_bwd on the packed layout."""

    def __init__(self, d_model: int, nhead: int, dim_feedforward: int = FFN_DIM, dropout: float = 0.0,
                 layer_norm_eps: float = 1e-5, batch_first: bool = True):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("chadavit_b200: dropout / drop_path > 0 is not supported (the reference never enables it)")
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.0, batch_first=batch_first)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(0.0)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm_first = False
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.dropout1 = nn.Dropout(0.0)
        self.dropout2 = nn.Dropout(0.0)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("TransformerEncoderLayer is executed by ChAdaViT on the packed token buffer; call the backbone")


class TokenLearner(nn.Module):
    """Image to patch embedding (chada_vit.py:118-134): parameter container + geometry."""

    def __init__(self, img_size: int = 224, patch_size: int = 16, in_chans: int = 1, embed_dim: int = 768):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = (img_size // patch_size) * (img_size // patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("TokenLearner is fused into the packed tokenizer kernel; call the backbone")


class _Saved:
    """Activations kept for the backward pass of one backbone call."""
    __slots__ = ("lay", "patches", "blocks", "x_last", "fin_idx", "fin_mean", "fin_rstd", "interp", "hw", "tail")


class _BackboneFn(torch.autograd.Function):
    """Autograd bridge: ONE node for the whole backbone so ``loss.backward()`` of the unchanged callers works."""

    @staticmethod
    def forward(ctx, module: "ChAdaViT", x: torch.Tensor, counts: Tuple[int, ...], *params):
        out, saved = module._forward_impl(x, counts, save=True)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        m: ChAdaViT = ctx.module
        gflat = torch.zeros_like(m.arena.fp32)
        m._backward_impl(ctx.saved, dout.contiguous().float(), gflat)
        ctx.saved = None
        grads = tuple(m.arena.g32(n, gflat) if p.requires_grad else None for n, p in zip(m.arena.names, m.arena.params))
        return (None, None, None) + grads


class ChAdaViT(nn.Module):
    """Channel Adaptive Vision Transformer (drop-in for chada_vit.py:136-330)."""

    def __init__(self, img_size=[224], in_chans=1, embed_dim=192, patch_size=16, num_classes=0, depth=12, num_heads=12,
                 drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, return_all_tokens=True, max_number_channels=10, **kwargs):
        super().__init__()
        if drop_rate != 0.0 or drop_path_rate != 0.0:
            raise NotImplementedError("chadavit_b200: drop_rate / drop_path_rate > 0 is not supported (reference default is 0)")
        if in_chans != 1:
            raise ValueError("ChAdaViT tokenises one channel at a time (in_chans must be 1)")
        if embed_dim % num_heads or embed_dim % 32:
            raise ValueError("embed_dim must be a multiple of 32 and divisible by num_heads")
        if embed_dim // num_heads not in (16, 32, 64, 96, 128):
            raise ValueError(f"unsupported head_dim {embed_dim // num_heads}: the sm_100a attention kernels support 16/32/64/96/128")
        self.num_features = self.embed_dim = embed_dim
        self.max_channels = max_number_channels
        self.num_heads = num_heads
        self.depth = depth
        self.token_learner = TokenLearner(img_size=img_size[0], patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.token_learner.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.channel_token = nn.Parameter(torch.zeros(1, self.max_channels, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=0.0)
        self.blocks = nn.ModuleList([TransformerEncoderLayer(embed_dim, num_heads, FFN_DIM, 0.0) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.return_all_tokens = return_all_tokens
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.cls_token, std=.02)
        trunc_normal_(self.channel_token, std=.02)
        self.apply(self._init_weights)          # chada_vit.py:171-183 (Conv2d / in_proj keep torch defaults)
        self._arena: Optional[ParamArena] = None
        self._interp: Dict[tuple, torch.Tensor] = {}
        # "bf16": tensor-core operands rounded to bf16 (the product path).  "split3": the fp32 parity run of the linear layers
        # (no-grad forward only): every operand of the qkv / out / linear1 / linear2 products is split into bf16 hi + lo and the
        # product runs over a three times longer K (x_hi w_hi + x_hi w_lo + x_lo w_hi, the k-NN kernel's operand layout,
        # csrc/knn.cu), i.e. to ~2^-16 — what remains of bf16 is the attention kernel's q / k / v / P and the patch embedding.
        self.linear_precision = "bf16"

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ------------------------------------------------------------------ parameter plumbing
    @property
    def arena(self) -> ParamArena:
        if self._arena is None:
            self._arena = ParamArena(self)
        return self._arena

    def _ready(self) -> ParamArena:
        a = self.arena
        a.ensure()
        if a.fp32.device.type != "cuda":
            raise RuntimeError("chadavit_b200.ChAdaViT runs on CUDA (sm_100a) only: move the module to the GPU; there is no CPU fallback")
        a.refresh_bf16()
        return a

    def _layout(self, counts: Sequence[int], npatch: int, device) -> ops.PackedLayout:
        return ops.get_layout(counts, npatch, device, PAD_CHANNELS)     # process-wide LRU shared by student and teacher

    def _interp_matrix(self, hp: int, wp: int, H: int, W: int, device) -> torch.Tensor:
        """(hp*wp, N0) fp32 map equal to the reference's bicubic resize of the patch position grid
        (add_pos_encoding_per_channel, chada_vit.py:201-217, incl. the DINO +0.1 trick); built once per geometry."""
        key = (hp, wp, str(device))
        if key not in self._interp:
            N0 = self.pos_embed.shape[2] - 1
            s = int(math.sqrt(N0))
            eye = torch.eye(N0).reshape(1, N0, s, s)
            w0, h0 = H // self.token_learner.patch_size + 0.1, W // self.token_learner.patch_size + 0.1
            m = F.interpolate(eye, scale_factor=(w0 / math.sqrt(N0), h0 / math.sqrt(N0)), mode="bicubic")
            assert int(w0) == m.shape[-2] and int(h0) == m.shape[-1]
            self._interp[key] = m.reshape(N0, -1).t().contiguous().to(device)
        return self._interp[key]

    # ------------------------------------------------------------------ public API (reference signatures)
    def forward(self, x: torch.Tensor, index: int, list_num_channels: List[List[int]]) -> torch.Tensor:
        counts = tuple(int(c) for c in list_num_channels[index])
        if not x.is_cuda:
            raise RuntimeError("chadavit_b200.ChAdaViT needs CUDA inputs (no CPU fallback)")
        self._ready()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.arena.params):
            return _BackboneFn.apply(self, x, counts, *self.arena.params)
        return self._forward_impl(x, counts, save=False)[0]

    @torch.no_grad()
    def get_last_selfattention(self, x: torch.Tensor) -> torch.Tensor:
        """Attention probabilities of the last block, (ΣC, num_heads, 1+N, 1+N) fp32 (chada_vit.py:313-320; main_attn.py reads
        ``attentions[0, :, 0, 1:]``).  The reference tokenises with ``list_num_channels=[1], max_channels=1``: every channel image
        is a sequence of its own, and the channel token is added only by a model built with ``max_number_channels == 1``."""
        if not x.is_cuda:
            raise RuntimeError("chadavit_b200.ChAdaViT needs CUDA inputs (no CPU fallback)")
        a = self._ready()
        counts = (1,) * x.shape[0]
        h, lay, _, _ = self._tokenize(x, counts, chan_rule=1)
        pre = None
        for i in range(self.depth - 1):
            h, _, pre = self._block_fwd(i, h, lay, False, pre)
        p = f"blocks.{self.depth - 1}."
        if pre is not None:
            u = pre[0]
        else:
            u, _, _, _ = ops.layernorm_fwd(h, a.v32(p + "norm1.weight"), a.v32(p + "norm1.bias"), self.blocks[-1].norm1.eps, save_stats=False)
        qkv = ops.gemm(u, a.v16(p + "self_attn.in_proj_weight"), bias=a.v32(p + "self_attn.in_proj_bias"))
        return ops.attn_probs(qkv, lay, self.num_heads)

    def _tokenize(self, x: torch.Tensor, counts: Sequence[int], chan_rule: int = PAD_CHANNELS):
        """Packed tokens of a ragged batch: (tokens fp32 [T,D], layout, bf16 patches, interp map or None).  The channel token is
        added iff ``self.max_channels == chan_rule`` (chada_vit.py:248: ``max_channels == self.max_channels`` with the
        ``max_channels`` argument 10 in forward and 1 in get_last_selfattention)."""
        a = self.arena
        D, P = self.embed_dim, self.token_learner.patch_size
        G, _, H, W = x.shape
        hp, wp = H // P, W // P
        N, N0 = hp * wp, self.pos_embed.shape[2] - 1
        lay = self._layout(counts, N, x.device)
        pos = a.v32("pos_embed").view(N0 + 1, D)
        interp = None
        if N == N0 and H == W:
            pos_patch = pos[1:]
        else:
            interp = self._interp_matrix(hp, wp, H, W, x.device)
            pos_patch = ops.small_matmul_f32(interp, pos[1:])
        chan = a.v32("channel_token").view(self.max_channels, D) if self.max_channels == chan_rule else None
        tok, patches = ops.tokenize_fwd(x, lay, P, a.v16("token_learner.proj.weight").view(D, P * P), a.v32("token_learner.proj.bias"),
                                        pos_patch, pos[0], a.v32("cls_token").view(D), chan)
        return tok, lay, patches, interp

    # ------------------------------------------------------------------ forward on the packed layout
    def _forward_impl(self, x: torch.Tensor, counts: Sequence[int], save: bool):
        a = self.arena
        H, W = x.shape[2], x.shape[3]
        tok, lay, patches, interp = self._tokenize(x, counts)
        blocks = []
        h, pre = tok, None          # pre = (u, mean, rstd) of this block's norm1(x) when the previous block already produced it
        if self.linear_precision != "bf16":               # fp32 parity run of the linear layers (see __init__)
            if self.linear_precision != "split3" or save:
                raise ValueError("linear_precision must be 'bf16' or 'split3' (split3: no-grad forward only)")
            for i in range(self.depth):
                h = self._block_fwd_split3(i, h, lay)
            idx = lay.non_cls_rows() if self.return_all_tokens else lay.cu[:-1]
            return ops.layernorm_fwd(h, a.v32("norm.weight"), a.v32("norm.bias"), self.norm.eps, in_idx=idx, out_bf16=False, out_f32=True,
                                     save_stats=False)[1], None
        # Only the CLS row of the final norm is returned (chada_vit.py:289) and every op behind the last attention is row-wise:
        # the last block then runs its attention for the CLS queries only and the rest on the B CLS rows (_tail_fwd).
        tail = ops.cls_tail_ok(self.depth, self.return_all_tokens, self.embed_dim // self.num_heads)
        for i in range(self.depth - 1 if tail else self.depth):
            h, sv, pre = self._block_fwd(i, h, lay, save, pre)
            blocks.append(sv)
        if tail:
            h, sv = self._tail_fwd(h, lay, save, pre)      # h: fp32 [B, D], the CLS rows only
            blocks.append(sv)
            idx = None
        else:
            idx = lay.non_cls_rows() if self.return_all_tokens else lay.cu[:-1]
        _, out, mean, rstd = ops.layernorm_fwd(h, a.v32("norm.weight"), a.v32("norm.bias"), self.norm.eps, in_idx=idx, out_bf16=False,
                                               out_f32=True, save_stats=save)
        if not save:
            return out, None
        s = _Saved()
        s.lay, s.patches, s.blocks, s.x_last, s.fin_idx, s.fin_mean, s.fin_rstd, s.interp, s.hw = lay, patches, blocks, h, idx, mean, rstd, interp, (H, W)
        s.tail = tail
        return out, s

    def _block_fwd_split3(self, i: int, x: torch.Tensor, lay: ops.PackedLayout) -> torch.Tensor:
        """One encoder block (chada_vit.py:95-116) with fp32-grade linear layers: activations and weights enter cb_gemm_bf16 as
        [hi | hi | lo] x [hi | lo | hi] bf16 operands over K = 3 x in_features (cb_split_bf16x3).  The parity run that shows what
        bf16 operand rounding of the linear layers costs; ~3x the tensor work and unfused, never the training path."""
        a, pre = self.arena, f"blocks.{i}."
        eps = self.blocks[i].norm1.eps
        g1, b1 = a.v32(pre + "norm1.weight"), a.v32(pre + "norm1.bias")
        R = ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32

        def lin(inp32, wname, bname, **kw):
            A, _ = ops.split_bf16x3(inp32.contiguous(), role_b=False, normalize=False)
            W, _ = ops.split_bf16x3(a.v32(pre + wname).contiguous(), role_b=True, normalize=False)
            return ops.gemm(A, W, bias=a.v32(pre + bname), **kw)
        u32 = ops.layernorm_fwd(x, g1, b1, eps, out_bf16=False, out_f32=True, save_stats=False)[1]
        qkv = lin(u32, "self_attn.in_proj_weight", "self_attn.in_proj_bias")                      # bf16: the attention kernel's input
        att, _ = ops.attn_fwd(qkv, lay, self.num_heads, need_lse=False)
        z1 = lin(att.float(), "self_attn.out_proj.weight", "self_attn.out_proj.bias", aux=x, flags=R)
        y32 = ops.layernorm_fwd(z1, g1, b1, eps, out_bf16=False, out_f32=True, save_stats=False)[1]
        hid = lin(y32, "linear1.weight", "linear1.bias", flags=ops.EPI_OUT_F32).relu_()
        z2 = lin(hid, "linear2.weight", "linear2.bias", aux=y32, flags=R)
        return ops.layernorm_fwd(z2, a.v32(pre + "norm2.weight"), a.v32(pre + "norm2.bias"), self.blocks[i].norm2.eps, out_bf16=False,
                                 out_f32=True, save_stats=False)[1]

    def _tail_fwd(self, x: torch.Tensor, lay: ops.PackedLayout, save: bool, pre_u=None):
        """The last encoder block when only the CLS embedding leaves the backbone.  x fp32 [T, D] -> x' fp32 [B, D] (CLS rows).
        qkv is still formed for every token (all of them are keys / values of the CLS query, chada_vit.py:105-111); the
        attention runs for the CLS queries only (cb_attn_cls_fwd), and the out-projection, norm1, feed-forward and norm2
        (chada_vit.py:99-100, :113-116) on the B CLS rows: identical results, the other rows were never read by anyone."""
        i = self.depth - 1
        a, pre = self.arena, f"blocks.{i}."
        eps = self.blocks[i].norm1.eps
        g1, b1 = a.v32(pre + "norm1.weight"), a.v32(pre + "norm1.bias")
        R = ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32
        if pre_u is not None:
            u, m1a, r1a = pre_u
        else:
            u, _, m1a, r1a = ops.layernorm_fwd(x, g1, b1, eps, save_stats=save)
        qkv = ops.gemm(u, a.v16(pre + "self_attn.in_proj_weight"), bias=a.v32(pre + "self_attn.in_proj_bias"))
        att, lse = ops.attn_cls_fwd(qkv, lay, self.num_heads)
        xc = x.index_select(0, lay.cls_rows64())
        z1 = ops.gemm(att, a.v16(pre + "self_attn.out_proj.weight"), bias=a.v32(pre + "self_attn.out_proj.bias"), aux=xc, flags=R)
        y, y32, m1b, r1b = ops.layernorm_fwd(z1, g1, b1, eps, out_f32=True, save_stats=save)
        hid = ops.gemm(y, a.v16(pre + "linear1.weight"), bias=a.v32(pre + "linear1.bias"), flags=ops.EPI_RELU)
        z2 = ops.gemm(hid, a.v16(pre + "linear2.weight"), bias=a.v32(pre + "linear2.bias"), aux=y32, flags=R)
        _, out, m2, r2 = ops.layernorm_fwd(z2, a.v32(pre + "norm2.weight"), a.v32(pre + "norm2.bias"), self.blocks[i].norm2.eps,
                                           out_bf16=False, out_f32=True, save_stats=save)
        if not save:
            return out, None
        return out, (x, u, m1a, r1a, qkv, att, lse, z1, m1b, r1b, y, hid, z2, m2, r2)

    def _block_fwd(self, i: int, x: torch.Tensor, lay: ops.PackedLayout, save: bool, pre_u=None):
        """x fp32 [T,D] -> x' fp32 [T,D].  The residual stream and every LayerNorm input stay fp32; bf16 is used only for
        tensor-core operands (u, qkv, att, y, hid).  norm2 of this block and norm1 of the next run as ONE kernel
        (ops.layernorm2_fwd): the third return value hands (u, mean, rstd) of the next block's norm1 to its caller."""
        a, pre = self.arena, f"blocks.{i}."
        eps = self.blocks[i].norm1.eps
        g1, b1 = a.v32(pre + "norm1.weight"), a.v32(pre + "norm1.bias")
        R = ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32
        if pre_u is not None:
            u, m1a, r1a = pre_u
        else:
            u, _, m1a, r1a = ops.layernorm_fwd(x, g1, b1, eps, save_stats=save)
        qkv = ops.gemm(u, a.v16(pre + "self_attn.in_proj_weight"), bias=a.v32(pre + "self_attn.in_proj_bias"))
        att, lse = ops.attn_fwd(qkv, lay, self.num_heads, need_lse=save)
        if ops.gemm_ln_ok(x.shape[0], x.shape[1], x.shape[1]):   # out-projection + residual + second use of norm1 in ONE kernel
            z1, y, y32, m1b, r1b = ops.gemm_ln_fwd(att, a.v16(pre + "self_attn.out_proj.weight"), a.v32(pre + "self_attn.out_proj.bias"), x,
                                                    g1, b1, eps, keep_z=save, save_stats=save)
        else:
            z1 = ops.gemm(att, a.v16(pre + "self_attn.out_proj.weight"), bias=a.v32(pre + "self_attn.out_proj.bias"), aux=x, flags=R)
            y, y32, m1b, r1b = ops.layernorm_fwd(z1, g1, b1, eps, out_f32=True, save_stats=save)
        if ops.ffn_fused_ok(x.shape[1], FFN_DIM):   # linear1 -> ReLU -> linear2 -> +residual in one kernel; hidden stored only if saved
            z2, hid, bits = ops.ffn_fwd(y, a.v16(pre + "linear1.weight"), a.v32(pre + "linear1.bias"), a.v16(pre + "linear2.weight"),
                                        a.v32(pre + "linear2.bias"), y32, save_hidden=save, save_mask_bits=save and _MASK_BITS)
        else:
            bits = None
            hid = ops.gemm(y, a.v16(pre + "linear1.weight"), bias=a.v32(pre + "linear1.bias"), flags=ops.EPI_RELU)
            z2 = ops.gemm(hid, a.v16(pre + "linear2.weight"), bias=a.v32(pre + "linear2.bias"), aux=y32, flags=R)
        nxt = None
        if i + 1 < self.depth and x.shape[1] in (64, 128, 192, 256):
            npre = f"blocks.{i + 1}."
            out, un, m2, r2, mn, rn = ops.layernorm2_fwd(z2, a.v32(pre + "norm2.weight"), a.v32(pre + "norm2.bias"), self.blocks[i].norm2.eps,
                                                         a.v32(npre + "norm1.weight"), a.v32(npre + "norm1.bias"), self.blocks[i + 1].norm1.eps,
                                                         save_stats=save)
            nxt = (un, mn, rn)
        else:
            _, out, m2, r2 = ops.layernorm_fwd(z2, a.v32(pre + "norm2.weight"), a.v32(pre + "norm2.bias"), self.blocks[i].norm2.eps,
                                               out_bf16=False, out_f32=True, save_stats=save)
        if not save:
            return out, None, nxt
        return out, (x, u, m1a, r1a, qkv, att, lse, z1, m1b, r1b, y, hid, z2, m2, r2, bits), nxt

    # ------------------------------------------------------------------ backward on the packed layout
    def _backward_impl(self, s: _Saved, dout: torch.Tensor, gflat: torch.Tensor, block_done=None) -> None:
        """Accumulates parameter gradients of one backbone call into the fp32 arena-shaped buffer ``gflat``.
        ``block_done(i)`` (optional) is called when every gradient of block i (and of everything behind it) has been queued;
        ``block_done(-1)`` after the tokenizer — the training engine hangs its bucketed gradient all-reduce on it."""
        a = self.arena
        g = lambda n: a.g32(n, gflat)  # noqa: E731
        D, P = self.embed_dim, self.token_learner.patch_size
        lay = s.lay
        # final norm (+ CLS / all-token gather): rows not selected get zero gradient
        dx, _ = ops.layernorm_bwd(dout, s.x_last, a.v32("norm.weight"), s.fin_mean, s.fin_rstd, dgamma=g("norm.weight"),
                                  dbeta=g("norm.bias"), idx=s.fin_idx)
        dx16 = None
        for i in reversed(range(self.depth)):
            if s.tail and i == self.depth - 1:
                dx = self._tail_bwd(s.blocks[i], dx, lay, gflat)
            else:
                dx, dx16 = self._block_bwd(i, s.blocks[i], dx, lay, gflat, last=(i == 0))
            s.blocks[i] = None
            if block_done is not None:
                block_done(i)
        N0 = self.pos_embed.shape[2] - 1
        dpos = g("pos_embed").view(N0 + 1, D)
        dpos_patch = dpos[1:] if s.interp is None else torch.zeros(lay.npatch, D, device=dx16.device, dtype=torch.float32)
        dchan = g("channel_token").view(self.max_channels, D) if self.max_channels == PAD_CHANNELS else None
        ops.tokenize_bwd(dx16, s.patches, lay, dw_pe=g("token_learner.proj.weight").view(D, P * P), db_pe=g("token_learner.proj.bias"),
                         dpos_patch=dpos_patch, dpos0=dpos[0], dcls_tok=g("cls_token").view(D), dchan_tok=dchan)
        if s.interp is not None:  # pos_embed receives gradient through the bicubic resize (SURVEY.md §8c probe)
            ops.small_matmul_f32(s.interp, dpos_patch, trans_a=True, out=dpos[1:], accumulate=True)
        if block_done is not None:
            block_done(-1)

    def _tail_bwd(self, sv, dxo: torch.Tensor, lay: ops.PackedLayout, gflat: torch.Tensor) -> torch.Tensor:
        """Backward of _tail_fwd: dxo fp32 [B, D] (gradient of the B CLS rows) -> dx fp32 [T, D].  Up to d(attention output)
        everything lives on the CLS rows (the gradient of every other row is exactly zero); cb_attn_cls_bwd turns it into the
        dense dqkv (dK / dV of every token), from where the block continues as _block_bwd does."""
        i = self.depth - 1
        a, pre = self.arena, f"blocks.{i}."
        g = lambda n: a.g32(pre + n, gflat)  # noqa: E731
        x, u, m1a, r1a, qkv, att, lse, z1, m1b, r1b, y, hid, z2, m2, r2 = sv
        T, D = x.shape
        Bc = z2.shape[0]
        A = ops.EPI_ATOMIC
        cs = ops.gemm_rowsum_ok(D)
        dz2, dz2h = ops.layernorm_bwd(dxo, z2, a.v32(pre + "norm2.weight"), m2, r2, dgamma=g("norm2.weight"), dbeta=g("norm2.bias"),
                                      dcolsum=g("linear2.bias"), want_bf16=True)
        ops.gemm(dz2h, hid, a_mn=True, b_mn=True, flags=A, out=g("linear2.weight"), k_splits=ops.splitk_wave(Bc, D, FFN_DIM))
        dh = ops.gemm(dz2h, a.v16(pre + "linear2.weight"), b_mn=True, aux=hid, flags=ops.EPI_RELU_MASK, colsum=None if cs else g("linear1.bias"))
        ops.gemm(dh, y, a_mn=True, b_mn=True, flags=A, out=g("linear1.weight"), k_splits=ops.splitk_wave(Bc, FFN_DIM, D),
                 colsum=g("linear1.bias") if cs else None)
        dy = ops.gemm(dh, a.v16(pre + "linear1.weight"), b_mn=True, aux=dz2, flags=ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32)
        dz1, dz1h = ops.layernorm_bwd(dy, z1, a.v32(pre + "norm1.weight"), m1b, r1b, dgamma=g("norm1.weight"), dbeta=g("norm1.bias"),
                                      dcolsum=g("self_attn.out_proj.bias"), want_bf16=True)
        ops.gemm(dz1h, att, a_mn=True, b_mn=True, flags=A, out=g("self_attn.out_proj.weight"), k_splits=ops.splitk_wave(Bc, D, D))
        datt = ops.gemm(dz1h, a.v16(pre + "self_attn.out_proj.weight"), b_mn=True)
        dqkv = ops.attn_cls_bwd(datt, qkv, att, lse, lay, self.num_heads)
        if not cs:
            ops.colsum(dqkv, g("self_attn.in_proj_bias"))
        ops.gemm(dqkv, u, a_mn=True, b_mn=True, flags=A, out=g("self_attn.in_proj_weight"), k_splits=ops.splitk_wave(T, 3 * D, D),
                 colsum=g("self_attn.in_proj_bias") if cs else None)
        du = ops.gemm(dqkv, a.v16(pre + "self_attn.in_proj_weight"), b_mn=True, flags=ops.EPI_OUT_F32)
        dx, _ = ops.layernorm_bwd(du, x, a.v32(pre + "norm1.weight"), m1a, r1a, dgamma=g("norm1.weight"), dbeta=g("norm1.bias"))
        dx.index_add_(0, lay.cls_rows64(), dz1)      # residual into z1 = x + attn: non-zero on the CLS rows only
        return dx

    def _block_bwd(self, i: int, sv, dxo: torch.Tensor, lay: ops.PackedLayout, gflat: torch.Tensor, last: bool):
        """dxo fp32 [T,D] -> (dx fp32, dx bf16 if last).  Gradient residual stream fp32; bf16 only for MMA operands."""
        a, pre = self.arena, f"blocks.{i}."
        g = lambda n: a.g32(pre + n, gflat)  # noqa: E731
        x, u, m1a, r1a, qkv, att, lse, z1, m1b, r1b, y, hid, z2, m2, r2, bits = sv
        T, D = x.shape
        A = ops.EPI_ATOMIC
        sk = lambda m, n: ops.splitk_wave(T, m, n)  # noqa: E731   (split-K of dW[m, n]: one full wave of CTAs)
        # x' = LN2(z2), z2 = y + relu(y W1^T + b1) W2^T + b2
        dz2, dz2h = ops.layernorm_bwd(dxo, z2, a.v32(pre + "norm2.weight"), m2, r2, dgamma=g("norm2.weight"), dbeta=g("norm2.bias"),
                                      dcolsum=g("linear2.bias"), want_bf16=True)
        ops.gemm(dz2h, hid, a_mn=True, b_mn=True, flags=A, out=g("linear2.weight"), k_splits=sk(D, FFN_DIM))
        # d(hidden) = (dz2 W2) o (hidden > 0): the mask comes as 1 bit per unit from the fused forward (hid itself: 16x the bytes)
        # bias gradients db = dY.sum(0) ride on the weight-gradient products dW = dY^T X where the kernel supports it
        # (ops.gemm_rowsum_ok: tensor-pipe sums of the A operand); otherwise the d(hidden) epilogue / a column-sum pass form them
        cs = ops.gemm_rowsum_ok(D)
        fused = bits is not None and cs and ops.ffn_bwd_fused_ok(D, FFN_DIM)
        if fused:   # d(hidden) and dy in one kernel: d(hidden) is written once (for dW1) and not read back for dy
            dy, dh = ops.ffn_bwd(dz2h, a.v16(pre + "linear2.weight"), a.v16(pre + "linear1.weight"), bits, dz2)
        else:
            dh = ops.gemm(dz2h, a.v16(pre + "linear2.weight"), b_mn=True, aux=hid if bits is None else bits,
                          flags=ops.EPI_RELU_MASK | (0 if bits is None else ops.EPI_MASK_BITS), colsum=None if cs else g("linear1.bias"))
        ops.gemm(dh, y, a_mn=True, b_mn=True, flags=A, out=g("linear1.weight"), k_splits=sk(FFN_DIM, D),
                 colsum=g("linear1.bias") if cs else None)
        if not fused:
            dy = ops.gemm(dh, a.v16(pre + "linear1.weight"), b_mn=True, aux=dz2, flags=ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32)
        # y = LN1(z1), z1 = x + att Wo^T + bo      (second use of norm1: gradients accumulate)
        dz1, dz1h = ops.layernorm_bwd(dy, z1, a.v32(pre + "norm1.weight"), m1b, r1b, dgamma=g("norm1.weight"), dbeta=g("norm1.bias"),
                                      dcolsum=g("self_attn.out_proj.bias"), want_bf16=True)
        ops.gemm(dz1h, att, a_mn=True, b_mn=True, flags=A, out=g("self_attn.out_proj.weight"), k_splits=sk(D, D))
        datt = ops.gemm(dz1h, a.v16(pre + "self_attn.out_proj.weight"), b_mn=True)
        dqkv = ops.attn_bwd(datt, qkv, att, lse, lay, self.num_heads)
        if not cs:
            ops.colsum(dqkv, g("self_attn.in_proj_bias"))
        ops.gemm(dqkv, u, a_mn=True, b_mn=True, flags=A, out=g("self_attn.in_proj_weight"), k_splits=sk(3 * D, D),
                 colsum=g("self_attn.in_proj_bias") if cs else None)
        du = ops.gemm(dqkv, a.v16(pre + "self_attn.in_proj_weight"), b_mn=True, flags=ops.EPI_OUT_F32)
        # u = LN1(x) (first use) ; dx = dLN1(du) + dz1 (residual into z1)
        return ops.layernorm_bwd(du, x, a.v32(pre + "norm1.weight"), m1a, r1a, dgamma=g("norm1.weight"), dbeta=g("norm1.bias"), dres=dz1,
                                 want_f32=not last, want_bf16=last)


def chada_vit(**kwargs):
    """Factory with the reference's semantics (chada_vit.py:333-339): depth 12, 2 heads, final LayerNorm eps 1e-6."""
    return ChAdaViT(patch_size=kwargs['patch_size'], embed_dim=kwargs['embed_dim'], depth=12, num_heads=2,
                    norm_layer=partial(nn.LayerNorm, eps=1e-6), return_all_tokens=kwargs['return_all_tokens'],
                    max_number_channels=kwargs['max_number_channels'])


def vit_channels(method, *args, **kwargs):
    """src/backbones/vit/__init__.py:57-59."""
    return chada_vit(**kwargs)
