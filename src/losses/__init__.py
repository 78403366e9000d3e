from .dino import DINOLoss  # noqa: F401
