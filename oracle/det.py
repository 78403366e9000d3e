"""TEST INFRASTRUCTURE — deterministic, RNG-free synthetic tensors.

Weights and pixels used by the golden fixtures are generated from a 64-bit integer
hash (splitmix64 finaliser) of (seed, element index), so the *same* values can be
regenerated bit-exactly on any machine with numpy: in the build container when the
reference modules produce the golden outputs (tests/golden/make_golden.py), and on
the GPU box when the CUDA path and the oracle are compared against them.  Nothing
relies on torch's RNG or on matching the reference's random init.
"""
from __future__ import annotations

import zlib

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def det_uniform(shape, seed: int, scale: float = 1.0, offset: float = 0.0) -> np.ndarray:
    """float32 array, uniform in [offset-scale, offset+scale), a pure function of (shape, seed)."""
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed & 0xFFFFFFFF) * np.uint64(0x100000001B3) + np.uint64(0x51ED27))
        h = _splitmix64(idx ^ key)
    u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)  # 24 bits -> exact in fp32
    out = ((u * 2.0 - 1.0) * scale + offset).astype(np.float32)
    return out.reshape(shape)


def name_seed(name: str, salt: int = 0) -> int:
    return (zlib.crc32(name.encode()) + 7919 * salt) & 0xFFFFFFFF


def param_scale(name: str, shape=()) -> tuple[float, float]:
    """(scale, offset) used for a state-dict entry, chosen so attention is neither flat nor one-hot (attention scores
    have a standard deviation of about 1.5 for LayerNorm-ed inputs, like a trained model) and LN is not the identity."""
    if name.endswith("in_proj_weight"):
        return float((4.5 / shape[1]) ** 0.5), 0.0
    if name.endswith("weight_g"):
        return 0.0, 1.0  # reference pins weight_g to 1 (src/methods/dino.py:81)
    if name.endswith("weight_v"):
        return 0.10, 0.0
    if "norm" in name and name.endswith("weight"):
        return 0.25, 1.0
    if "norm" in name and name.endswith("bias"):
        return 0.10, 0.0
    if name.endswith("running_var"):
        return 0.20, 1.0
    if name.endswith("running_mean"):
        return 0.10, 0.0
    if name.endswith("num_batches_tracked"):
        return 0.0, 0.0
    if len(shape) == 1 and name.endswith("weight"):   # BatchNorm1d of DINOHead(use_bn=True)
        return 0.25, 1.0
    if name.endswith("bias"):
        return 0.05, 0.0
    if name in ("cls_token", "channel_token", "pos_embed"):
        return 0.50, 0.0
    if name.endswith("proj.weight"):  # token_learner conv
        return 0.10, 0.0
    return 0.08, 0.0


def det_state_dict(shapes: dict, salt: int = 0) -> dict:
    """name -> float32 ndarray for every (name, shape) in ``shapes`` (insertion order kept)."""
    out = {}
    for name, shape in shapes.items():
        sc, off = param_scale(name, tuple(shape))
        out[name] = det_uniform(tuple(shape), name_seed(name, salt), sc, off)
    return out


def det_pixels(total_channels: int, h: int, w: int, seed: int) -> np.ndarray:
    """(ΣC,1,H,W) float32 pixels in [-1,1): one_channel_collate_fn layout (src/data/channels_strategies.py:31-85)."""
    return det_uniform((total_channels, 1, h, w), 0xC0FFEE + seed, 1.0, 0.0)
