"""Overlay tree with the reference's module paths: every leaf module re-exports the B200-native class of
the same name from ``chadavit_b200``, so that the reference's own imports

    from src.backbones.vit.chada_vit import ChAdaViT          (src/methods/base.py:36, src/methods/linear.py:46)
    from src.losses.dino import DINOLoss                      (src/methods/dino.py:27)
    from src.utils.momentum import MomentumUpdater, initialize_momentum_params   (src/methods/base.py:57)
    from src.utils.lars import LARS                           (src/methods/base.py:56)

resolve to the CUDA path.  ``tools/install_overlay.py <reference checkout>`` copies the LEAF files (never an ``__init__.py``)
over the same-named files of a reference checkout; the package ``__init__`` files here exist only so that this tree is
importable on its own (tests, notebooks).  Nothing in this tree computes anything.
"""
