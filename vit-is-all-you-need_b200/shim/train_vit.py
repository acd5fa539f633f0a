"""Drop-in for `from train_vit import ViTConfig, ViT` (train_titok.py:8, train_vit_vqgan.py:8): the classes of
train_vit.py:16-53 with the fused patch embedding."""
from b200vit.modules import ViT, ViTClassifier, ViTConfig  # noqa: F401
