"""LARS behind the reference's optimizer interface (src/utils/lars.py:21-167) for the autograd drop-in path
(``DINO.configure_optimizers`` with ``optimizer.name: lars``, the pre-training yaml's choice, src/methods/base.py:67-72).

Same constructor, ``param_groups`` / ``state_dict`` behaviour of a ``torch.optim.Optimizer``; ``step()`` runs the flat-arena
kernels (``cb_param_norms`` + ``cb_lars_step``, csrc/optim.cu) — one norm pass and one update pass per network instead of
two ``torch.norm`` host synchronisations and ~8 elementwise launches per parameter.  Every parameter must belong to a
chadavit_b200 module (their parameters are views of one fp32 arena, arena.py); there is no per-tensor fallback.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch.optim.optimizer import Optimizer

from .. import ops
from ..arena import ParamArena, arena_of_param

_HYPER = ("lr", "momentum", "dampening", "nesterov", "eta", "eps", "clip_lr")


class LARS(Optimizer):
    def __init__(self, params, lr, momentum=0, dampening=0, weight_decay=0, nesterov=False, eta=1e-3, eps=1e-8, clip_lr=False,
                 exclude_bias_n_norm=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov, eta=eta, eps=eps,
                        clip_lr=clip_lr, exclude_bias_n_norm=exclude_bias_n_norm)
        super().__init__(params, defaults)
        self._ws: Dict[int, dict] = {}

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("nesterov", False)

    def _workspace(self, arena: ParamArena) -> dict:
        ws = self._ws.get(id(arena))
        if ws is None or ws["buf"].device != arena.fp32.device or ws["buf"].numel() != arena.numel:
            dev = arena.fp32.device
            ws = self._ws[id(arena)] = {
                "buf": torch.zeros_like(arena.fp32), "partial": torch.empty(arena.numel // 32, device=dev, dtype=torch.float32),
                "norms": torch.zeros(len(arena.names) * 3, device=dev, dtype=torch.float32),
                "noclip": torch.zeros(len(arena.names), dtype=torch.uint8, device=dev), "flags": {}, "stepped": set()}
        return ws

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # (arena, hyper-parameters, non-zero weight decay) -> [(name, decays, adapts)]: one launch each
        plan: Dict[Tuple, List[Tuple[str, bool, bool]]] = {}
        arenas: Dict[int, ParamArena] = {}
        for group in self.param_groups:
            hyper = tuple(group[k] for k in _HYPER)
            for p in group["params"]:
                if p.grad is None:
                    continue
                arena, name = arena_of_param(p)
                if arena.fp32.device.type != "cuda":
                    raise RuntimeError("chadavit_b200.LARS runs on CUDA only (no CPU fallback)")
                arenas[id(arena)] = arena
                adapts = p.ndim != 1 or not group["exclude_bias_n_norm"]                       # lars.py:136
                wd = float(group["weight_decay"])
                plan.setdefault((id(arena), hyper, wd if wd != 0 else None), []).append((name, wd != 0, adapts))
        # one kernel launch handles one non-zero weight-decay value: fold the wd == 0 groups of an arena into a launch that has one
        merged: Dict[Tuple, List[Tuple[str, bool, bool]]] = {}
        for (aid, hyper, wd), items in plan.items():
            if wd is None:
                host = next((k for k in plan if k[0] == aid and k[1] == hyper and k[2] is not None), None)
                merged.setdefault(host or (aid, hyper, 0.0), []).extend(items)
            else:
                merged.setdefault((aid, hyper, wd), []).extend(items)
        for aid, arena in arenas.items():
            arena.ensure()
            ws = self._workspace(arena)
            g = arena.ensure_grad()
            g.zero_()
            views, grads = [], []
            for n, p in zip(arena.names, arena.params):
                if p.grad is not None:
                    views.append(arena.g32(n, g))
                    grads.append(p.grad.to(torch.float32))
            torch._foreach_copy_(views, grads)
            start, seg_of = arena.segment_maps()
            ops.param_norms(arena.fp32, g, start, ws["noclip"], ws["partial"], ws["norms"], grad_scale=1.0, clip=0.0)
            for (aid2, hyper, wd), items in merged.items():
                if aid2 != aid:
                    continue
                h = dict(zip(_HYPER, hyper))
                first = tuple(n for n, _, _ in items if n not in ws["stepped"]) if (h["momentum"] != 0 and h["dampening"] != 0) else ()
                key = (tuple(items), first)
                fl = ws["flags"].get(key)
                if fl is None:
                    host = torch.full((arena.numel,), 2, dtype=torch.uint8)                    # everything else: untouched
                    for n, decays, adapts in items:
                        off, cnt, _ = arena.offsets[n]
                        host[off:off + cnt] = (1 if decays else 0) | (4 if adapts else 0) | (8 if n in first else 0)
                    if len(ws["flags"]) > 16:
                        ws["flags"].clear()
                    fl = ws["flags"][key] = host.to(arena.fp32.device)
                ops.lars_step(arena.fp32, g, ws["buf"], fl, seg_of, ws["norms"], lr=h["lr"], momentum=h["momentum"],
                              dampening=h["dampening"], nesterov=h["nesterov"], weight_decay=wd or 0.0, eta=h["eta"], eps=h["eps"],
                              clip_lr=h["clip_lr"], p_bf16=arena.bf16)
                ws["stepped"].update(n for n, _, _ in items)
            arena.mark_dirty()
            arena._bf16_key = (arena.manual_version, sum(p._version for p in arena.params))    # shadows were refreshed by the kernel
        return loss
