from .dino import DINO, DINOHead  # noqa: F401

METHODS = {"dino": DINO}
from .extract import extract_features  # noqa: F401
