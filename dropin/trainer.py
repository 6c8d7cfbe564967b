"""Drop-in for the reference's `trainer` module: `from trainer import Trainer` (main.py:5)."""
from uegan_b200.trainer import Trainer, ImagePool, init_weights  # noqa: F401
