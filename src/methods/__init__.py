from .dino import DINO, DINOHead  # noqa: F401

METHODS = {"dino": DINO}
