"""chadavit_b200 — B200 (sm_100a) native hot path for ChAda-ViT + DINO behind the reference's Python API.

    from chadavit_b200.backbones import ChAdaViT, chada_vit, vit_channels
    from chadavit_b200.methods import DINOHead, DINOTrainer
    from chadavit_b200.losses import DINOLoss
    from chadavit_b200.utils.momentum import MomentumUpdater, initialize_momentum_params

Compute goes through libchadavit_b200.so (C ABI in include/chadavit_b200.h); there is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
