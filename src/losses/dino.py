"""Drop-in for src/losses/dino.py of nicoboou/chadavit (``DINOLoss``: fused temperature softmax / centring / multi-crop
cross-entropy kernel + centre EMA), implemented by chadavit_b200."""
from chadavit_b200.losses.dino import DINOLoss  # noqa: F401

__all__ = ["DINOLoss"]
