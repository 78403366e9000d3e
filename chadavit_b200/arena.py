"""Flat parameter arenas: every parameter of a module is a view into ONE fp32 buffer (plus a bf16 shadow for the
tensor-core kernels and an fp32 gradient arena).  This is what lets the teacher EMA (src/utils/momentum.py:73-74),
the optimizer step and the fp32->bf16 refresh each be a single launch, and what a bucketed NCCL all-reduce runs over.

Parameters stay ordinary fp32 ``nn.Parameter`` objects with the reference's names, so ``state_dict`` /
``load_state_dict`` / ``parameters()`` order are unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import nn

ALIGN = 64  # elements: 256 B in fp32, 128 B in bf16 (TMA needs 16 B)


class ParamArena:
    def __init__(self, module: nn.Module):
        self.module = module
        self.names: List[str] = []
        self.params: List[nn.Parameter] = []
        self.offsets: Dict[str, Tuple[int, int, torch.Size]] = {}
        off = 0
        for name, p in module.named_parameters():
            self.names.append(name)
            self.params.append(p)
            self.offsets[name] = (off, p.numel(), p.shape)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.fp32: torch.Tensor = None  # type: ignore
        self.bf16: torch.Tensor = None  # type: ignore
        self.grad: torch.Tensor = None  # type: ignore  (allocated on demand by the training engine)
        self._bf16_key = None
        self.manual_version = 0
        self._flatten()

    # ------------------------------------------------------------------ layout
    def _flatten(self) -> None:
        dev = self.params[0].device
        flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for name, p in zip(self.names, self.params):
                off, n, shape = self.offsets[name]
                flat[off:off + n].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = flat[off:off + n].view(shape)
                p._cb_arena = (self, name)        # lets an optimizer handed bare parameters find their arena (utils/lars.py)
        self.fp32 = flat
        self.bf16 = torch.empty(self.numel, device=dev, dtype=torch.bfloat16) if dev.type == "cuda" else None
        self.grad = None
        self._bf16_key = None
        self._views32, self._views16 = {}, {}     # per-name views of the two arenas (rebuilt with the arenas)
        self._viewsg_of, self._viewsg = None, {}

    def ensure(self) -> None:
        """Re-flatten if some code replaced ``p.data`` (e.g. ``.to()``, the reference's EMA ``mp.data = ...``)."""
        base = self.fp32.data_ptr()
        for name, p in zip(self.names, self.params):
            if p.data_ptr() != base + 4 * self.offsets[name][0] or p.dtype != torch.float32:
                self._flatten()
                return

    def mark_dirty(self) -> None:
        """Call after writing the fp32 arena directly (fused optimizer / EMA kernels bypass autograd versions)."""
        self.manual_version += 1

    def refresh_bf16(self, force: bool = False) -> None:
        from . import ops
        key = (self.manual_version, sum(p._version for p in self.params))
        if force or key != self._bf16_key:
            ops.cast_bf16(self.fp32, self.bf16)
            self._bf16_key = key

    # ------------------------------------------------------------------ views
    # The views are cached: a training step asks for ~800 of them, and slice + view cost ~5 us each on the host.
    def v32(self, name: str) -> torch.Tensor:
        t = self._views32.get(name)
        if t is None:
            off, n, shape = self.offsets[name]
            t = self._views32[name] = self.fp32[off:off + n].view(shape)
        return t

    def v16(self, name: str) -> torch.Tensor:
        t = self._views16.get(name)
        if t is None:
            off, n, shape = self.offsets[name]
            t = self._views16[name] = self.bf16[off:off + n].view(shape)
        return t

    def g32(self, name: str, grad_flat: torch.Tensor) -> torch.Tensor:
        """View of parameter `name` inside an arena-shaped gradient buffer.  Cached only for the arena's OWN gradient buffer (the
        engine's persistent one), keyed by object identity: a view can never alias memory it was not made for, and temporary
        buffers (tests, the autograd bridge) are not kept alive by a cache."""
        off, n, shape = self.offsets[name]
        if grad_flat is not self.grad or grad_flat is None:
            return grad_flat[off:off + n].view(shape)
        if self._viewsg_of is not grad_flat:
            self._viewsg_of, self._viewsg = grad_flat, {}
        t = self._viewsg.get(name)
        if t is None:
            t = self._viewsg[name] = grad_flat[off:off + n].view(shape)
        return t

    def segment_maps(self):
        """(seg_start_block int32 [P+1], seg_of_block int32 [numel/64]) on the arena's device: which parameter owns each
        64-element block (alignment padding belongs to the parameter in front of it) — for the per-parameter norm kernels."""
        m = getattr(self, "_segmaps", None)
        if m is None or m[0].device != self.fp32.device:
            starts = [self.offsets[n][0] // ALIGN for n in self.names] + [self.numel // ALIGN]
            start = torch.tensor(starts, dtype=torch.int32)
            counts = start[1:] - start[:-1]
            of = torch.repeat_interleave(torch.arange(len(self.names), dtype=torch.int32), counts.long())
            m = self._segmaps = (start.to(self.fp32.device), of.to(self.fp32.device))
        return m

    def ensure_grad(self) -> torch.Tensor:
        if self.grad is None or self.grad.device != self.fp32.device:
            self.grad = torch.zeros_like(self.fp32)
        return self.grad


def arena_of_param(p: torch.Tensor):
    """(arena, parameter name) of a parameter owned by a chadavit_b200 module; raises for foreign parameters (no fallback)."""
    ref = getattr(p, "_cb_arena", None)
    if ref is None:
        raise RuntimeError("parameter does not belong to a chadavit_b200 module (ChAdaViT / DINOHead): the fused optimizers work on "
                           "the modules' flat parameter arenas and have no per-tensor fallback")
    ref[0].ensure()
    return ref
