from .vit import vit_channels  # noqa: F401
