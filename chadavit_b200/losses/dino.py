"""DINOLoss behind the reference's interface (src/losses/dino.py:27-118), computed by ONE fused CUDA kernel
(temperature softmax of the centred teacher, student log-softmax, multi-crop cross-entropy and d(loss)/d(student)) plus
the centre column-sum / EMA kernels.  Same constructor, attributes (``epoch``, ``center`` buffer, ``teacher_temp_schedule``),
call signature and update order (the loss uses the OLD centre; the centre is updated afterwards: SURVEY.md Q13)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops


class _DINOLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, center, V, student_temp, teacher_temp):
        loss, d32, _ = ops.dino_loss_fwd_bwd(student.detach().contiguous().float(), teacher.detach().contiguous().float(), center, V,
                                             student_temp, teacher_temp, want_f32=True, want_bf16=False)
        ctx.save_for_backward(d32)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d32,) = ctx.saved_tensors
        return d32 * g, None, None, None, None, None


class DINOLoss(nn.Module):
    def __init__(self, num_prototypes: int, warmup_teacher_temp: float, teacher_temp: float, warmup_teacher_temp_epochs: float,
                 num_epochs: int, student_temp: float = 0.1, num_large_crops: int = 2, center_momentum: float = 0.9):
        super().__init__()
        self.epoch = 0
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.num_large_crops = num_large_crops
        self.register_buffer("center", torch.zeros(1, num_prototypes))
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(num_epochs - warmup_teacher_temp_epochs) * teacher_temp,
        ))

    def forward(self, student_output: torch.Tensor, teacher_output: torch.Tensor) -> torch.Tensor:
        if not student_output.is_cuda:
            raise RuntimeError("chadavit_b200.DINOLoss needs CUDA tensors (no CPU fallback)")
        temp = float(self.teacher_temp_schedule[self.epoch])
        loss = _DINOLossFn.apply(student_output, teacher_output, self.center.view(-1), self.num_large_crops, self.student_temp, temp)
        self.update_center(teacher_output)
        return loss

    @torch.no_grad()
    def update_center(self, teacher_output: torch.Tensor, run_collective=None):
        """sum -> all_reduce(SUM) -> / world / rows -> EMA (src/losses/dino.py:111-118); the EMA is done in place.
        ``run_collective(fn)`` (optional) runs the all-reduce + EMA tail, e.g. on a side stream that the caller joins before the
        centre is read again (the training engine overlaps it with the backward pass)."""
        t = teacher_output.detach().contiguous().float()
        ws = getattr(self, "_sum_ws", None)      # persistent workspace: it may be read on a side stream after this call returns
        if ws is None or ws.device != t.device or ws.numel() != t.shape[1]:
            ws = self._sum_ws = torch.empty(t.shape[1], device=t.device, dtype=torch.float32)
        batch_sum = ops.colsum_f32(t, out=ws)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        world = dist.get_world_size() if multi else 1

        def tail():
            if multi:
                dist.all_reduce(batch_sum)
            ops.center_ema(self.center.view(-1), batch_sum, 1.0 / (world * t.shape[0]), self.center_momentum)
        if run_collective is not None and multi:
            run_collective(tail)
        else:
            tail()
