"""Drop-in for the reference's `losses` module: `from losses import PerceptualLoss, GANLoss, MultiscaleRecLoss`
(trainer.py:9) and `from losses import PerceptualLoss, TVLoss` (tester.py:9)."""
from uegan_b200.losses import *  # noqa: F401,F403
from uegan_b200.losses import GANLoss, MultiscaleRecLoss, PerceptualLoss, TVLoss, VGG19_relu  # noqa: F401
