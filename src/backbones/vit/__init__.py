"""src/backbones/vit/__init__.py:57-59 of the reference: the ``vit_channels`` factory used by BaseMethod."""
from .chada_vit import chada_vit as default_chada_vit


def vit_channels(method, *args, **kwargs):
    return default_chada_vit(*args, **kwargs)


__all__ = ["vit_channels"]
