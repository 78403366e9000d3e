"""Drop-in for src/backbones/vit/chada_vit.py of nicoboou/chadavit: same module path, same public names
(``ChAdaViT``, ``TransformerEncoderLayer``, ``TokenLearner``, ``chada_vit``), implemented by chadavit_b200 (sm_100a kernels
behind the C ABI of include/chadavit_b200.h).  ``isinstance(backbone, ChAdaViT)`` checks of the reference
(src/methods/base.py:526, src/methods/linear.py:389) hold because this IS the class the factory builds."""
from chadavit_b200.backbones.chada_vit import (  # noqa: F401
    ChAdaViT,
    TokenLearner,
    TransformerEncoderLayer,
    chada_vit,
    trunc_normal_,
)

__all__ = ["ChAdaViT", "TransformerEncoderLayer", "TokenLearner", "chada_vit"]
