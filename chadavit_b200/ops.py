"""Thin Python wrappers over the C ABI: torch tensors in, raw device pointers across the boundary.

PyTorch is used only for device memory and streams.  Every function launches on the current CUDA stream and is
asynchronous.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import collections
import math
import os

import numpy as np
import torch

from . import _lib

EPI_RELU, EPI_RESIDUAL, EPI_RELU_MASK, EPI_OUT_F32, EPI_ATOMIC, EPI_RESIDUAL_F32, EPI_MASK_BITS = 1, 2, 4, 8, 16, 64, 128
bf16 = torch.bfloat16


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("chadavit_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


# Raw handle of the current stream of the current device.  torch.cuda.current_stream() resolves the device index through several
# Python layers (incl. os.environ lookups) and builds a Stream object: ~15 us per call, 430 calls per training step — it was
# 23 % of the host time of a step (tools/host_profile.py), and the host is only just ahead of the GPU on fresh ragged batches.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


# kernels launched per C-ABI call (for bench.py's gpu_launches claim; memsets are not counted)
_KERNELS_PER_CALL = {"cb_tokenize_fwd": 2, "cb_tokenize_bwd": 2, "cb_attn_varlen_bwd": 3, "cb_sync_check": 0, "cb_param_norms": 2}

# Optional per-kernel-class device timing (bench.py): PROFILE[name] = [n_launches, flops, [(start_evt, end_evt), ...], bytes]
PROFILE = None


def _call(name: str, *args, work: float = 0.0, nbytes: float = 0.0, pkey: Optional[str] = None) -> None:
    """pkey: profiling class when one entry point dispatches to different kernels (cb_ffn_fwd with / without the hidden store)."""
    lib = _lib.load()
    _lib.launch_count += _KERNELS_PER_CALL.get(name, 1)
    if PROFILE is not None and (pkey or name) in PROFILE:
        rec = PROFILE[pkey or name]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(getattr(lib, name)(*args), name)
        e1.record()
        rec[0] += 1
        rec[1] += work
        rec[2].append((e0, e1))
        if len(rec) > 3:
            rec[3] += nbytes
        return
    _lib.check(getattr(lib, name)(*args), name)


def sync_check() -> None:
    _call("cb_sync_check", _stream())


# ------------------------------------------------------------------------------------------------ GEMM
def gemm(A: torch.Tensor, B: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, bias: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, flags: int = 0, out: Optional[torch.Tensor] = None, alpha: float = 1.0,
         k_splits: int = 1, colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M,N] (+)= alpha * op(A) op(B)^T.  a_mn/b_mn: the operand is stored [K, M] / [K, N] (MN-major)."""
    assert A.dtype == bf16 and B.dtype == bf16 and A.dim() == 2 and B.dim() == 2
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    N, Kb = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
    assert K == Kb, f"K mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=torch.float32 if flags & (EPI_OUT_F32 | EPI_ATOMIC) else bf16)
    assert out.shape == (M, N) and out.stride(1) == 1
    _call("cb_gemm_bf16", _p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(out), out.stride(0), M, N, K,
          _p(bias), _p(aux), aux.stride(0) if aux is not None else 0, flags, float(alpha), k_splits, _p(colsum), _stream(), work=2.0 * M * N * K,
          nbytes=2.0 * (M * K + N * K) + float(M) * N * out.element_size() +
          (0.0 if aux is None else float(M) * N / 8 if flags & EPI_MASK_BITS else float(M) * N * aux.element_size()))
    return out


_GEMM_LN = os.environ.get("CB_NO_GEMM_LN", "") != "1"   # A/B switch (read once): LayerNorm fused into the out-projection epilogue


def gemm_ln_ok(M: int, N: int, K: int) -> bool:
    return _GEMM_LN and N == 192 and K <= 192 and K % 8 == 0 and M >= 512


def gemm_ln_fwd(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], resid: Optional[torch.Tensor], gamma: torch.Tensor,
                beta: torch.Tensor, eps: float, *, keep_z: bool, out_f32: bool = True, save_stats: bool = True):
    """z = A W^T + bias + resid ; y = LayerNorm(z) in one kernel (N = 192).  Returns (z fp32 | None, y bf16, y fp32 | None, mean, rstd)."""
    M, K = A.shape
    N = W.shape[0]
    assert A.dtype == bf16 and W.dtype == bf16 and A.stride(1) == 1 and W.stride(1) == 1 and W.shape[1] == K
    assert resid is None or (resid.dtype == torch.float32 and resid.shape == (M, N) and resid.stride(1) == 1)
    dev = A.device
    z = torch.empty(M, N, device=dev, dtype=torch.float32) if keep_z else None
    y = torch.empty(M, N, device=dev, dtype=bf16)
    y32 = torch.empty(M, N, device=dev, dtype=torch.float32) if out_f32 else None
    mean = torch.empty(M, device=dev, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(M, device=dev, dtype=torch.float32) if save_stats else None
    _call("cb_gemm_ln_fwd", _p(A), A.stride(0), _p(W), W.stride(0), _p(bias), _p(resid), resid.stride(0) if resid is not None else 0,
          _p(gamma), _p(beta), float(eps), _p(z), _p(y), _p(y32), _p(mean), _p(rstd), M, N, K, _stream(), work=2.0 * M * N * K,
          pkey="cb_gemm_bf16", nbytes=float(M) * (K * 2 + N * (4 + 2 + (4 if out_f32 else 0) + (4 if keep_z else 0))) + 2.0 * N * K)
    return z, y, y32, mean, rstd


_GEMM_ROWSUM = os.environ.get("CB_NO_GEMM_ROWSUM", "") != "1"   # A/B switch (read once): bias gradients on the tensor pipe of the dW products


def gemm_rowsum_ok(N: int) -> bool:
    """True when a weight-gradient product gemm(dY, X, a_mn=True, b_mn=True, flags=EPI_ATOMIC) with X of width N can also
    deliver the bias gradient dY.sum(0) through ``colsum`` (the 192-wide column tile has the spare accumulator columns)."""
    return _GEMM_ROWSUM and N % 192 == 0 and N % 256 != 0


def ffn_fused_ok(D: int, F: int) -> bool:
    return D == 192 and F % 64 == 0 and F <= 2048


def ffn_fwd(y: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, resid: torch.Tensor, *,
            save_hidden: bool, save_mask_bits: Optional[bool] = None, kernel: int = 0):
    """z2 = resid + relu(y W1^T + b1) W2^T + b2 in one kernel (D = 192).  Returns (z2 fp32, hidden bf16 | None), or — when the
    ``save_mask_bits`` keyword is given at all — (z2, hidden | None, bits | None): the ReLU mask as bits, uint32 [F/32, ld]
    (for gemm(..., flags=EPI_RELU_MASK | EPI_MASK_BITS)).  ``kernel``: 0 = the library's choice, 1 / 3 force one (tests, A/B)."""
    T, D = y.shape
    F = w1.shape[0]
    assert y.dtype == bf16 and w1.dtype == bf16 and w2.dtype == bf16 and resid.dtype == torch.float32
    assert y.is_contiguous() and w1.is_contiguous() and w2.is_contiguous() and resid.is_contiguous() and w2.shape == (D, F)
    z2 = torch.empty(T, D, device=y.device, dtype=torch.float32)
    hid = torch.empty(T, F, device=y.device, dtype=bf16) if save_hidden else None
    bits = torch.empty(F // 32, (T + 31) // 32 * 32, device=y.device, dtype=torch.int32) if save_mask_bits else None
    _call("cb_ffn_fwd", _p(y), _p(w1), _p(b1), _p(w2), _p(b2), _p(resid), _p(z2), _p(hid), _p(bits), bits.shape[1] if bits is not None else 0,
          T, D, F, int(kernel), _stream(), work=4.0 * T * D * F, pkey="cb_ffn_fwd" if save_hidden else "cb_ffn_fwd:nostore",
          nbytes=float(T) * (D * 2 + D * 8 + (F * 2 if save_hidden else 0) + (F / 8 if save_mask_bits else 0)) + 4.0 * D * F)
    return (z2, hid) if save_mask_bits is None else (z2, hid, bits)


_FFN_BWD_FUSED = os.environ.get("CB_NO_FFN_BWD", "") != "1"   # A/B switch (read once): fused d(hidden) + dy kernel


def ffn_bwd_fused_ok(D: int, F: int) -> bool:
    return _FFN_BWD_FUSED and ffn_fused_ok(D, F)


def ffn_bwd(dz2h: torch.Tensor, w2: torch.Tensor, w1: torch.Tensor, bits: torch.Tensor, dz2: torch.Tensor):
    """dh = (dz2 W2) o (hidden > 0) (bf16 [T, F], stored for dW1) and dy = dh W1 + dz2 (fp32 [T, D]) in one kernel; ``bits`` is the
    ReLU mask written by ffn_fwd(save_mask_bits=True).  Returns (dy, dh)."""
    T, D = dz2h.shape
    F = w1.shape[0]
    assert dz2h.dtype == bf16 and w1.dtype == bf16 and w2.dtype == bf16 and dz2.dtype == torch.float32 and bits.dtype == torch.int32
    assert dz2h.is_contiguous() and w1.is_contiguous() and w2.is_contiguous() and dz2.is_contiguous() and bits.is_contiguous()
    assert w2.shape == (D, F) and w1.shape == (F, D) and dz2.shape == (T, D) and bits.shape[0] == F // 32 and bits.shape[1] >= T
    dy = torch.empty(T, D, device=dz2h.device, dtype=torch.float32)
    dh = torch.empty(T, F, device=dz2h.device, dtype=bf16)
    _call("cb_ffn_bwd", _p(dz2h), _p(w2), _p(w1), _p(bits), bits.shape[1], _p(dz2), _p(dy), _p(dh), T, D, F, _stream(),
          work=4.0 * T * D * F, nbytes=float(T) * (D * 2 + D * 8 + F * 2 + F / 8) + 4.0 * D * F)
    return dy, dh


def splitk_for(K: int, tiles: int, target_ctas: int = 148) -> int:
    """Number of K splits so that a weight-gradient GEMM with `tiles` output tiles fills the GPU."""
    kb = (K + 63) // 64
    return max(1, min(kb, (target_ctas + tiles - 1) // tiles))


_SPLITK_LEGACY = os.environ.get("CB_SPLITK_LEGACY", "") == "1"


def gemm_tiles(M: int, N: int) -> int:
    """Output tiles of cb_gemm_bf16 for an [M, N] result: 128-row tiles x the column tile the kernel picks (gemm.cu, gemm_run:
    192 when N is a multiple of 192 but not of 256, else 256 for N >= 256, else 128)."""
    bn = 192 if (N % 192 == 0 and N % 256 != 0) else (256 if N >= 256 else 128)
    return ((M + 127) // 128) * ((N + bn - 1) // bn)


def splitk_wave(K: int, M: int, N: int, n_ctas: int = 148) -> int:
    """Split-K factor of a weight-gradient product dW[M, N] = dY^T X over K tokens such that tiles x splits fills ONE wave of the
    persistent kernel (<= n_ctas CTAs, as close to it as the tile count allows).  The products are HBM-bound: with the tile
    counts the call sites used to estimate (128-wide column tiles; the kernel uses 192 / 256) dW1 ran 80 CTAs on 148 SMs, dWqkv 75,
    dWo 74, and dW2 a full wave plus a 12-CTA tail: 145 / 172 / 58 / 29 us per 137 k tokens against 107 / 117 / 41 / 21 us with one
    full wave (profiles/r02_microbench_splitk.txt)."""
    kb = (K + 63) // 64
    if _SPLITK_LEGACY:      # A/B only: the round-1 estimate (128-wide column tiles, rounded up)
        return splitk_for(K, ((M + 127) // 128) * (N // 256 if N >= 2048 else (N + 127) // 128))
    return max(1, min(kb, n_ctas // max(1, gemm_tiles(M, N))))


# ------------------------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, in_idx: Optional[torch.Tensor] = None,
                  out_bf16: bool = True, out_f32: bool = False, save_stats: bool = True):
    """x fp32 [*, D] -> (y bf16 | None, y fp32 | None, mean, rstd)."""
    assert x.dtype == torch.float32
    D = x.shape[-1]
    rows = in_idx.numel() if in_idx is not None else x.shape[0]
    y = torch.empty(rows, D, device=x.device, dtype=bf16) if out_bf16 else None
    y32 = torch.empty(rows, D, device=x.device, dtype=torch.float32) if out_f32 else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    _call("cb_layernorm_fwd", _p(x), _p(in_idx), _p(gamma), _p(beta), _p(y), _p(y32), _p(mean), _p(rstd), rows, D, float(eps), _stream(),
          nbytes=float(rows) * D * (4 + (2 if out_bf16 else 0) + (4 if out_f32 else 0)))
    return y, y32, mean, rstd


def layernorm2_fwd(x: torch.Tensor, gamma_a, beta_a, eps_a: float, gamma_b, beta_b, eps_b: float, *, save_stats: bool = True):
    """y1 = LN_a(x) fp32, y2 = LN_b(y1) bf16 in one pass.  Returns (y1, y2, mean_a, rstd_a, mean_b, rstd_b)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    rows, D = x.shape
    y1 = torch.empty(rows, D, device=x.device, dtype=torch.float32)
    y2 = torch.empty(rows, D, device=x.device, dtype=bf16)
    st = [torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None for _ in range(4)]
    _call("cb_layernorm2_fwd", _p(x), _p(gamma_a), _p(beta_a), float(eps_a), _p(gamma_b), _p(beta_b), float(eps_b), _p(y1), _p(y2),
          _p(st[0]), _p(st[1]), _p(st[2]), _p(st[3]), rows, D, _stream(), nbytes=float(rows) * D * 10)
    return (y1, y2, *st)


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, *,
                  dgamma: torch.Tensor, dbeta: torch.Tensor, dcolsum: Optional[torch.Tensor] = None,
                  dres: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None, want_f32: bool = True,
                  want_bf16: bool = False):
    """dy, x, dres fp32.  Returns (dx fp32 | None, dx bf16 | None); with idx the untouched rows are zero."""
    assert dy.dtype == torch.float32 and x.dtype == torch.float32
    D = x.shape[-1]
    rows = dy.shape[0]
    mk = torch.zeros if idx is not None else torch.empty
    dx32 = mk(x.shape, device=x.device, dtype=torch.float32) if want_f32 else None
    dx16 = mk(x.shape, device=x.device, dtype=bf16) if want_bf16 else None
    _call("cb_layernorm_bwd", _p(dy), _p(x), _p(idx), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx32), _p(dx16), _p(dgamma),
          _p(dbeta), _p(dcolsum), rows, D, _stream(),
          nbytes=float(rows) * D * (8 + (4 if dres is not None else 0) + (4 if want_f32 else 0) + (2 if want_bf16 else 0)))
    return dx32, dx16


def colsum(x: torch.Tensor, out: torch.Tensor) -> None:
    _call("cb_colsum_bf16", _p(x), x.stride(0), _p(out), x.shape[0], x.shape[1], _stream())


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert src.dtype == torch.float32 and src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=bf16)
    _call("cb_cast_f32_bf16", _p(src), _p(dst), src.numel(), _stream())
    return dst


# ------------------------------------------------------------------------------------------------ varlen bookkeeping
class PackedLayout:
    """Host-side index bookkeeping for one ragged batch (bit-exact contract).

    counts[b] = C_b channels of image b.  Sequence b owns packed rows [cu[b], cu[b+1]) with 1 + C_b*N rows.
    Everything is derived from Python ints (list_num_channels) -> no device synchronisation.
    """

    def __init__(self, counts: Sequence[int], npatch: int, device, max_channels: int = 10):
        counts = [int(c) for c in counts]
        for c in counts:
            if c < 1 or c > max_channels:
                raise ValueError(f"number of channels per image must be in [1, {max_channels}], got {c}")
        self.counts = counts
        self.npatch = npatch
        self.B = len(counts)
        self.G = sum(counts)
        lens = np.array([1 + c * npatch for c in counts], dtype=np.int64)
        cu = np.zeros(self.B + 1, dtype=np.int32)
        cu[1:] = np.cumsum(lens)
        self.cu_host = cu
        self.T = int(cu[-1])
        self.max_seqlen = int(lens.max())
        self.sum_sq = float((lens.astype(np.float64) ** 2).sum())   # sum_b S_b^2: attention work
        chan_img = np.repeat(np.arange(self.B, dtype=np.int32), counts)
        chan_idx = np.concatenate([np.arange(c, dtype=np.int32) for c in counts])
        self.chan_img_host, self.chan_idx_host = chan_img, chan_idx
        self.device = device
        self.cu = torch.from_numpy(cu).to(device, non_blocking=True)
        self.chan_img = torch.from_numpy(chan_img).to(device, non_blocking=True)
        self.chan_idx = torch.from_numpy(chan_idx).to(device, non_blocking=True)
        # int64 copy of the CLS rows for torch gathers / scatters.  Made HERE, with the other index tensors: the training engine
        # creates its layouts on the compute stream before it forks the teacher / local-crop chains onto side streams, and a
        # tensor first uploaded inside one chain would be read by the others without a stream dependency.
        self._cls64_host = cu[:-1].astype(np.int64)
        self._cls64 = torch.from_numpy(self._cls64_host).to(device, non_blocking=True)
        self._work = {}
        self._work_host = {}

    def attn_work(self, nheads: int, tile: int = 128) -> torch.Tensor:
        """(n_work, 4) int32 {q_row0, seq_start, seq_end, head}, longest sequences first (LPT order)."""
        return self._schedule(nheads, tile, "list", 0)

    def attn_schedule(self, nheads: int, tile: int, kind: str, n_ctas: Optional[int] = None) -> torch.Tensor:
        """Work list of `attn_work` re-ordered for the persistent kernels, which give CTA c the slots c, c + G, c + 2G, ...:
        items are assigned longest-processing-time-first to the least loaded of the G CTAs and written back round-major,
        padded with empty slots (seq_end <= seq_start, skipped by the kernels).  Pure host arithmetic on the Python ints of
        list_num_channels (cb_attn_schedule, csrc/host.cu): no device synchronisation, ~50 us per schedule.
        kind = "fwd" (cost ~ kv sub-tiles x query tiles of the item) or "bwd" (cost ~ query tiles per kv tile)."""
        return self._schedule(nheads, tile, kind, n_ctas or _lib.load().cb_num_sms())

    def _schedule(self, nheads: int, tile: int, kind: str, G: int) -> torch.Tensor:
        key = (nheads, tile, kind, G)
        w = self._work.get(key)
        if w is None:
            import ctypes as C
            lib = _lib.load()
            n_items = int(nheads * ((self.cu_host[1:] - self.cu_host[:-1] + tile - 1) // tile).sum())
            cap = n_items if kind == "list" else (n_items // max(G, 1) + 8) * max(G, 1) + n_items // 4
            n_out = C.c_int(0)
            while True:
                host = np.zeros((cap, 4), dtype=np.int32)
                rc = lib.cb_attn_schedule(self.cu_host.ctypes.data, self.B, nheads, tile, {"list": 0, "fwd": 1, "bwd": 2}[kind], G,
                                          host.ctypes.data, cap, C.byref(n_out))
                if rc == 0:
                    break
                if n_out.value <= cap:
                    _lib.check(rc, "cb_attn_schedule")
                cap = n_out.value
            self._work_host[key] = host[:n_out.value]
            w = self._work[key] = torch.from_numpy(self._work_host[key]).to(self.device, non_blocking=True)
        return w

    def cls_rows64(self) -> torch.Tensor:
        """Packed row of every sequence's CLS token (its first row) as an int64 index for torch gathers / scatters."""
        return self._cls64

    def non_cls_rows(self) -> torch.Tensor:
        """Packed rows of all patch tokens in (b, c, p) order (return_all_tokens=True output order, chada_vit.py:283-287)."""
        if not hasattr(self, "_noncls"):
            rows = np.concatenate([np.arange(self.cu_host[b] + 1, self.cu_host[b + 1], dtype=np.int32) for b in range(self.B)])
            self._noncls = torch.from_numpy(rows).to(self.device, non_blocking=True)
        return self._noncls


# One layout cache for the whole process (student and teacher see the same batches): LRU, strong references.  A CUDA graph
# that baked a layout's device pointers keeps the layout alive itself (KEEPALIVE), so eviction here can never free memory a
# captured graph still reads.
_LAYOUTS: "collections.OrderedDict[tuple, PackedLayout]" = collections.OrderedDict()
LAYOUT_CACHE_SIZE = 32
KEEPALIVE: Optional[list] = None      # set to a list while a CUDA graph is being captured: every layout handed out is appended


def get_layout(counts: Sequence[int], npatch: int, device, max_channels: int = 10) -> PackedLayout:
    key = (tuple(counts), npatch, str(device), max_channels)
    lay = _LAYOUTS.get(key)
    if lay is None:
        lay = _LAYOUTS[key] = PackedLayout(counts, npatch, device, max_channels)
        while len(_LAYOUTS) > LAYOUT_CACHE_SIZE:
            _LAYOUTS.popitem(last=False)
    else:
        _LAYOUTS.move_to_end(key)
    if KEEPALIVE is not None:
        KEEPALIVE.append(lay)
    return lay


# ------------------------------------------------------------------------------------------------ tokenizer / attention
def tokenize_fwd(x: torch.Tensor, lay: PackedLayout, patch: int, w_pe_bf16: torch.Tensor, b_pe: torch.Tensor, pos_patch: torch.Tensor,
                 pos0: torch.Tensor, cls_tok: torch.Tensor, chan_tok: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    G, one, H, W = x.shape
    if one != 1:
        raise ValueError("ChAdaViT expects a (sum_channels, 1, H, W) tensor")
    if G != lay.G:
        raise ValueError(f"x has {G} channel images but list_num_channels sums to {lay.G}")
    if H % patch or W % patch:
        raise ValueError(f"image size {H}x{W} must be a multiple of the patch size {patch}")
    if x.dtype != torch.float32:
        x = x.float()
    if not (x.stride(3) == 1 and x.stride(2) == W and x.stride(0) == H * W):  # size-1 dim is stride-agnostic (channels_last)
        x = x.contiguous()
    D = w_pe_bf16.shape[0]
    patches = torch.empty(lay.T, patch * patch, device=x.device, dtype=bf16)
    tokens = torch.empty(lay.T, D, device=x.device, dtype=torch.float32)
    _call("cb_tokenize_fwd", _p(x), G, H, W, patch, _p(lay.cu), _p(lay.chan_img), lay.B, _p(w_pe_bf16), _p(b_pe), _p(pos_patch),
          _p(pos0), _p(cls_tok), _p(chan_tok), _p(patches), _p(tokens), lay.T, D, _stream())
    return tokens, patches


def tokenize_bwd(dtokens: torch.Tensor, patches: torch.Tensor, lay: PackedLayout, *, dw_pe, db_pe, dpos_patch, dpos0, dcls_tok, dchan_tok) -> None:
    T, D = dtokens.shape
    pe = patches.shape[1]
    ks = splitk_for(T, ((D + 127) // 128) * ((pe + 255) // 256))
    _call("cb_tokenize_bwd", _p(dtokens), _p(patches), _p(lay.cu), _p(lay.chan_img), _p(lay.chan_idx), lay.G, lay.B, lay.npatch, pe,
          T, D, _p(dw_pe), _p(db_pe), _p(dpos_patch), _p(dpos0), _p(dcls_tok), _p(dchan_tok), ks, _stream())


def attn_fwd(qkv: torch.Tensor, lay: PackedLayout, nheads: int, *, need_lse: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    T, D3 = qkv.shape
    D = D3 // 3
    d = D // nheads
    q_tile = 256                                             # work item = a pair of 128-row query tiles
    work = lay.attn_schedule(nheads, q_tile, "fwd")
    out = torch.empty(T, D, device=qkv.device, dtype=bf16)
    lse = torch.empty(nheads, T, device=qkv.device, dtype=torch.float32) if need_lse else None
    _call("cb_attn_varlen_fwd", _p(qkv), _p(work), work.shape[0], q_tile, _p(out), _p(lse), T, D, nheads, float(d) ** -0.5, _stream(),
          work=4.0 * D * lay.sum_sq, nbytes=8.0 * T * D)
    return out, lse


def attn_bwd(dout: torch.Tensor, qkv: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, lay: PackedLayout, nheads: int) -> torch.Tensor:
    """dqkv [T, 3D] bf16 from d(attention output) [T, D] bf16."""
    T, D3 = qkv.shape
    D = D3 // 3
    d = D // nheads
    work = lay.attn_schedule(nheads, 128, "bwd")
    delta = torch.empty(nheads, T, device=qkv.device, dtype=torch.float32)
    dq_acc = torch.empty(T, D, device=qkv.device, dtype=torch.float32)
    dqkv = torch.empty(T, D3, device=qkv.device, dtype=bf16)
    _call("cb_attn_varlen_bwd", _p(dout), _p(qkv), _p(out), _p(lse), _p(work), work.shape[0], _p(delta), _p(dq_acc), _p(dqkv), T, D,
          nheads, float(d) ** -0.5, _stream(), work=10.0 * D * lay.sum_sq, nbytes=18.0 * T * D)
    return dqkv


_CLS_TAIL = os.environ.get("CB_NO_CLS_TAIL", "") != "1"   # A/B switch (read once): CLS-only tail of the last block


def cls_tail_ok(depth: int, return_all_tokens: bool, head_dim: int) -> bool:
    """True when the last encoder block may run its attention for the CLS queries only and everything behind it on the CLS rows
    only: the backbone returns ``x[:, 0]`` (chada_vit.py:289), every op behind the last attention is row-wise."""
    return _CLS_TAIL and not return_all_tokens and depth >= 2 and head_dim in (16, 32, 64, 96, 128)


def attn_cls_fwd(qkv: torch.Tensor, lay: PackedLayout, nheads: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Attention output of the CLS query of every sequence: (out bf16 [B, D], lse fp32 [B, H])."""
    T, D3 = qkv.shape
    D = D3 // 3
    d = D // nheads
    assert qkv.dtype == bf16 and qkv.is_contiguous() and T == lay.T
    out = torch.empty(lay.B, D, device=qkv.device, dtype=bf16)
    lse = torch.empty(lay.B, nheads, device=qkv.device, dtype=torch.float32)
    _call("cb_attn_cls_fwd", _p(qkv), _p(lay.cu), lay.B, nheads, d, float(d) ** -0.5, _p(out), _p(lse), _stream(),
          work=4.0 * D * T, nbytes=4.0 * T * D)
    return out, lse


def attn_cls_bwd(dout: torch.Tensor, qkv: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, lay: PackedLayout, nheads: int) -> torch.Tensor:
    """dqkv bf16 [T, 3D] (every element written) from d(attention output) of the CLS rows, bf16 [B, D]."""
    T, D3 = qkv.shape
    D = D3 // 3
    d = D // nheads
    assert dout.dtype == bf16 and dout.is_contiguous() and dout.shape == (lay.B, D) and out.shape == (lay.B, D) and out.is_contiguous()
    assert lse.shape == (lay.B, nheads) and qkv.is_contiguous() and T == lay.T
    dqkv = torch.empty(T, D3, device=qkv.device, dtype=bf16)
    _call("cb_attn_cls_bwd", _p(dout), _p(qkv), _p(out), _p(lse), _p(lay.cu), lay.B, nheads, d, float(d) ** -0.5, _p(dqkv), _stream(),
          work=10.0 * D * T, nbytes=10.0 * T * D)
    return dqkv


def small_matmul_f32(A: torch.Tensor, B: torch.Tensor, *, trans_a: bool = False, out: Optional[torch.Tensor] = None,
                     accumulate: bool = False) -> torch.Tensor:
    """fp32 C (+)= op(A) @ B for the tiny pos-embed resize map only."""
    M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
    N = B.shape[1]
    assert B.shape[0] == K and A.is_contiguous() and B.is_contiguous()
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=torch.float32)
    _call("cb_small_matmul_f32", _p(A), _p(B), _p(out), M, N, K, int(trans_a), int(accumulate), _stream())
    return out


# ------------------------------------------------------------------------------------------------ DINO head / loss / EMA / optimizer
def gelu_fwd(pre: torch.Tensor) -> torch.Tensor:
    out = torch.empty(pre.shape, device=pre.device, dtype=bf16)
    _call("cb_gelu_fwd", _p(pre), _p(out), pre.numel(), _stream())
    return out


def gelu_bwd(dact: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    out = torch.empty(pre.shape, device=pre.device, dtype=bf16)
    _call("cb_gelu_bwd", _p(dact), _p(pre), _p(out), pre.numel(), _stream())
    return out


def bn_gelu_fwd(pre: torch.Tensor, gamma, beta, running_mean, running_var, momentum: float, eps: float, training: bool, save: bool):
    """BatchNorm1d + GELU (DINOHead(use_bn=True)): returns (act bf16, bn_out fp32 | None, save_mean | None, save_invstd | None)."""
    R, C = pre.shape
    act = torch.empty(R, C, device=pre.device, dtype=bf16)
    bn_out = torch.empty_like(pre) if save else None
    mean = torch.empty(C, device=pre.device, dtype=torch.float32) if save else None
    invstd = torch.empty(C, device=pre.device, dtype=torch.float32) if save else None
    _call("cb_bn_gelu_fwd", _p(pre), _p(gamma), _p(beta), _p(running_mean), _p(running_var), float(momentum), float(eps), int(bool(training)),
          _p(bn_out), _p(act), _p(mean), _p(invstd), R, C, _stream())
    return act, bn_out, mean, invstd


def bn_gelu_bwd(dact: torch.Tensor, bn_out, pre, gamma, mean, invstd, training: bool, dgamma, dbeta) -> torch.Tensor:
    R, C = pre.shape
    dpre = torch.empty(R, C, device=pre.device, dtype=bf16)
    _call("cb_bn_gelu_bwd", _p(dact), _p(bn_out), _p(pre), _p(gamma), _p(mean), _p(invstd), int(bool(training)), _p(dpre), _p(dgamma), _p(dbeta),
          R, C, _stream())
    return dpre


def l2norm_fwd(x: torch.Tensor, eps: float = 1e-12):
    rows, C = x.shape
    out = torch.empty(rows, C, device=x.device, dtype=bf16)
    inv = torch.empty(rows, device=x.device, dtype=torch.float32)
    _call("cb_l2norm_fwd", _p(x), _p(out), _p(inv), rows, C, float(eps), _stream())
    return out, inv


def l2norm_bwd(dy: torch.Tensor, x: torch.Tensor, inv: torch.Tensor) -> torch.Tensor:
    rows, C = x.shape
    dx = torch.empty(rows, C, device=x.device, dtype=bf16)
    _call("cb_l2norm_bwd", _p(dy), _p(x), _p(inv), _p(dx), rows, C, _stream())
    return dx


def weightnorm_fwd(v: torch.Tensor, g: torch.Tensor):
    K, C = v.shape
    w = torch.empty(K, C, device=v.device, dtype=bf16)
    inv = torch.empty(K, device=v.device, dtype=torch.float32)
    _call("cb_weightnorm_fwd", _p(v), _p(g), _p(w), _p(inv), K, C, _stream())
    return w, inv


def weightnorm_bwd(dw: torch.Tensor, v: torch.Tensor, g: torch.Tensor, inv: torch.Tensor, dv: torch.Tensor, dg: Optional[torch.Tensor]) -> None:
    K, C = v.shape
    _call("cb_weightnorm_bwd", _p(dw), _p(v), _p(g), _p(inv), _p(dv), _p(dg), K, C, _stream())


def dino_loss_fwd_bwd(student: torch.Tensor, teacher: torch.Tensor, center: torch.Tensor, V: int, student_temp: float, teacher_temp: float,
                      *, want_f32: bool = False, want_bf16: bool = True):
    """Returns (loss[1] fp32, dstudent fp32 | None, dstudent bf16 | None).  student (V*B,K), teacher (2*B,K) fp32."""
    assert student.dtype == torch.float32 and teacher.dtype == torch.float32 and student.is_contiguous() and teacher.is_contiguous()
    K = student.shape[1]
    B = teacher.shape[0] // 2
    if student.shape[0] != V * B or teacher.shape[0] != 2 * B or teacher.shape[1] != K:
        raise ValueError(f"DINOLoss: student {tuple(student.shape)} / teacher {tuple(teacher.shape)} do not match {V} student views of 2 teacher views")
    loss = torch.empty(1, device=student.device, dtype=torch.float32)
    d32 = torch.empty_like(student) if want_f32 else None
    d16 = torch.empty(student.shape, device=student.device, dtype=bf16) if want_bf16 else None
    _call("cb_dino_loss_fwd_bwd", _p(student), _p(teacher), _p(center), _p(loss), _p(d32), _p(d16), B, K, V, float(student_temp),
          float(teacher_temp), _stream())
    return loss, d32, d16


def colsum_f32(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty(x.shape[1], device=x.device, dtype=torch.float32)
    _call("cb_colsum_f32", _p(x), _p(out), x.shape[0], x.shape[1], _stream())
    return out


def center_ema(center: torch.Tensor, batch_sum: torch.Tensor, scale: float, momentum: float) -> None:
    _call("cb_dino_center_ema", _p(center), _p(batch_sum), float(scale), float(momentum), center.numel(), _stream())


def ema_update(momentum_flat: torch.Tensor, online_flat: torch.Tensor, tau: float, momentum_bf16: Optional[torch.Tensor] = None) -> None:
    assert momentum_flat.numel() == online_flat.numel()
    _call("cb_ema_update", _p(momentum_flat), _p(online_flat), _p(momentum_bf16), float(tau), momentum_flat.numel(), _stream())


def adamw_step(p, g, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1, step_late=None, flags=None, p_bf16=None,
               teacher=None, teacher_bf16=None, grad_scale=1.0, tau=1.0, dev_hyper=None) -> None:
    """step_late: step count of the elements flagged bit4 (parameters that started receiving gradients late; default = step).
    dev_hyper: optional device fp32[6] {lr, 1-beta1^step, sqrt(1-beta2^step), tau, 1-beta1^step_late, sqrt(1-beta2^step_late)}
    read by the kernel (CUDA-graph replay)."""
    _call("cb_adamw_step", _p(p), _p(g), _p(m), _p(v), _p(flags), _p(p_bf16), _p(teacher), _p(teacher_bf16), p.numel(), float(lr),
          float(beta1), float(beta2), float(eps), float(weight_decay), int(step), int(step if step_late is None else step_late),
          float(grad_scale), float(tau), _p(dev_hyper), _stream())


def param_norms(p, g, seg_start_block, seg_clip, partial, norms, *, grad_scale=1.0, clip=0.0) -> None:
    """norms[s] = (||p_s||, ||g_s*grad_scale*coef_s||, coef_s) per parameter of a flat arena (deterministic two-pass reduction);
    coef_s is the dino_clip_gradients coefficient (dino.py:249-261) where seg_clip[s] != 0 and clip > 0, else 1."""
    _call("cb_param_norms", _p(p), _p(g), _p(seg_start_block), _p(seg_clip), _p(partial), _p(norms), p.numel(),
          seg_start_block.numel() - 1, float(grad_scale), float(clip), _stream())


def scale_grads(g, seg_of_block, norms) -> None:
    """g *= coef of the owning parameter, in place (the clip coefficients computed by ``param_norms``)."""
    _call("cb_scale_grads", _p(g), _p(seg_of_block), _p(norms), g.numel(), _stream())


def lars_step(p, g, buf, flags, seg_of_block, norms, *, lr, momentum=0.0, dampening=0.0, nesterov=False, weight_decay=0.0, eta=1e-3,
              eps=1e-8, clip_lr=False, p_bf16=None, teacher=None, teacher_bf16=None, grad_scale=1.0, tau=1.0, dev_hyper=None) -> None:
    """LARS.step (src/utils/lars.py:113-167) over a flat arena + teacher EMA + bf16 shadows, one launch."""
    _call("cb_lars_step", _p(p), _p(g), _p(buf), _p(flags), _p(seg_of_block), _p(norms), _p(p_bf16), _p(teacher), _p(teacher_bf16),
          p.numel(), float(lr), float(momentum), float(dampening), int(bool(nesterov)), float(weight_decay), float(eta), float(eps),
          int(bool(clip_lr)), float(grad_scale), float(tau), _p(dev_hyper), _stream())


def attn_probs(qkv: torch.Tensor, lay: "PackedLayout", num_heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) per sequence and head as a dense fp32 (nseq, H, S_max, S_max) tensor (zero beyond a sequence's
    length) — the last block's attention map for get_last_selfattention (chada_vit.py:313-320)."""
    D = qkv.shape[1] // 3
    d = D // num_heads
    nseq = lay.cu.numel() - 1
    s_max = int(lay.max_seqlen)
    out = torch.empty(nseq, num_heads, s_max, s_max, device=qkv.device, dtype=torch.float32)
    _call("cb_attn_probs", _p(qkv), _p(lay.cu), nseq, num_heads, d, s_max, 1.0 / math.sqrt(d), _p(out), _stream())
    return out


def split_bf16x3(x: torch.Tensor, role_b: bool, normalize: bool, want_sqnorm: bool = False, pad_rows_to: int = 1):
    """fp32 [R, D] -> bf16 [R', 3D] operand ([hi|hi|lo] or, for the B side, [hi|lo|hi]) for an fp32-accurate tensor-core product;
    R' = R rounded up to ``pad_rows_to`` (extra rows zero).  Optionally L2-normalises rows first and returns ||row||^2."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    R, D = x.shape
    Rp = (R + pad_rows_to - 1) // pad_rows_to * pad_rows_to
    out = torch.empty(Rp, 3 * D, device=x.device, dtype=bf16)
    if Rp > R:
        out[R:].zero_()
    sq = torch.zeros(Rp, device=x.device, dtype=torch.float32) if want_sqnorm else None
    _call("cb_split_bf16x3", _p(x), _p(out), _p(sq), R, D, int(role_b), int(normalize), _stream())
    return out, sq


def inv_euclid_(dots: torch.Tensor, sq_a: torch.Tensor, sq_b: torch.Tensor, n_cols: int, eps: float) -> torch.Tensor:
    """dots[i, j] <- 1 / (sqrt(max(|a_i|^2 + |b_j|^2 - 2 dots[i, j], 0)) + eps) for the first n_cols columns, in place."""
    _call("cb_inv_euclid", _p(dots), _p(sq_a), _p(sq_b), dots.shape[0], n_cols, dots.stride(0), float(eps), _stream())
    return dots
