"""Ragged collate behind the reference's contract (src/data/channels_strategies.py:31-85).

The reference builds every crop batch with one ``unsqueeze`` per channel and a ``torch.cat`` over ΣC single-channel
tensors.  Here each image (``(C, H, W)``, channels contiguous) is copied ONCE, straight into its slot of one
preallocated — optionally pinned — ``(ΣC, 1, H, W)`` buffer per crop, which is what ``DINO.stage_batch`` then moves to
the GPU with a single asynchronous copy per crop.  The returned triple is bit-identical to the reference's:

    crop_lists            list (one per crop; a bare tensor when there is one crop) of ``(ΣC, 1, H, W)`` tensors,
                          channels of image 0, then image 1, ...
    batched_labels        ``torch.tensor(labels)``
    num_channels_lists    list (per crop) of lists (per image) of Python ints
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class OneChannelCollator:
    """Callable ``collate_fn``.  ``pin_memory=True`` returns page-locked crop buffers (they are allocated per call, so a
    DataLoader worker / prefetch queue can hold several batches at once); ``reuse=N`` instead cycles through N sets of
    pinned buffers per batch signature, for a training loop that stages each batch before asking for the next one."""

    def __init__(self, pin_memory: bool = False, reuse: int = 0):
        self.pin_memory = bool(pin_memory)
        self.reuse = int(reuse)
        self._pool: dict = {}
        self._turn: dict = {}

    def _buffers(self, shapes: Sequence[tuple], dtype: torch.dtype) -> List[torch.Tensor]:
        if self.reuse <= 0:
            return [torch.empty(s, dtype=dtype, pin_memory=self.pin_memory) for s in shapes]
        key = (tuple(shapes), dtype)
        sets = self._pool.setdefault(key, [])
        if len(sets) < self.reuse:
            sets.append([torch.empty(s, dtype=dtype, pin_memory=self.pin_memory) for s in shapes])
            return sets[-1]
        t = self._turn.get(key, 0)
        self._turn[key] = (t + 1) % self.reuse
        return sets[t]

    def __call__(self, batch):
        # datasets return (image, label) or (index, image, label): the last two entries count (channels_strategies.py:52)
        first = batch[0][-2:][0]
        num_crops = len(first) if isinstance(first, list) else 1
        num_channels_lists: List[List[int]] = [[] for _ in range(num_crops)]
        labels = []
        images: List[List[torch.Tensor]] = [[] for _ in range(num_crops)]
        for item in batch:
            image_list, label = item[-2:]
            if isinstance(image_list, torch.Tensor):
                image_list = [image_list]
            for i, crop in enumerate(image_list):
                num_channels_lists[i].append(int(crop.shape[0]))
                images[i].append(crop)
            labels.append(label)
        shapes, dtype = [], None
        for i in range(num_crops):
            h, w = images[i][0].shape[-2:]
            for im in images[i]:
                if im.dim() != 3 or tuple(im.shape[-2:]) != (h, w):
                    # same failure class as the reference's torch.cat over mismatched (1, H, W) channel images
                    raise RuntimeError(f"one_channel_collate_fn: crop {i} holds images of different sizes "
                                       f"({tuple(im.shape)} vs (*, {h}, {w}))")
            dt = images[i][0].dtype
            for im in images[i][1:]:
                dt = torch.promote_types(dt, im.dtype)                        # torch.cat's type promotion
            dtype = dt if dtype is None else dtype
            if dt != dtype:
                dtype = None
                break
            shapes.append((sum(num_channels_lists[i]), 1, h, w))
        if dtype is None:       # crops of different dtypes: one buffer set per crop
            bufs = []
            for i in range(num_crops):
                h, w = images[i][0].shape[-2:]
                dt = images[i][0].dtype
                for im in images[i][1:]:
                    dt = torch.promote_types(dt, im.dtype)
                bufs.append(torch.empty((sum(num_channels_lists[i]), 1, h, w), dtype=dt, pin_memory=self.pin_memory))
        else:
            bufs = self._buffers(shapes, dtype)
        for i in range(num_crops):
            off = 0
            flat = bufs[i].view(bufs[i].shape[0], bufs[i].shape[2], bufs[i].shape[3])
            for im in images[i]:
                c = im.shape[0]
                flat[off:off + c].copy_(im)                                       # one copy per image, not per channel
                off += c
        crop_lists = bufs[0] if num_crops == 1 else list(bufs)
        return crop_lists, torch.tensor(labels), num_channels_lists


_default = OneChannelCollator()


def one_channel_collate_fn(batch):
    """Drop-in for ``src.data.channels_strategies.one_channel_collate_fn`` (same triple, same values)."""
    return _default(batch)
