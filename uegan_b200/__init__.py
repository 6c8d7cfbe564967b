"""uegan_b200: B200-native (sm_100a) implementation of the UEGAN hot path behind the reference's Python API.

    from uegan_b200.models import Generator, Discriminator      # reference: models.py
    from uegan_b200.losses import PerceptualLoss, GANLoss, MultiscaleRecLoss   # reference: losses.py

All compute goes through libuegan_sm100.so (include/uegan_sm100.h); nothing here falls back to PyTorch ops."""
__version__ = "0.1.0"
