from .chada_vit import ChAdaViT, TokenLearner, TransformerEncoderLayer, chada_vit, vit_channels  # noqa: F401
