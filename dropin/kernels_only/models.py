"""Drop-in for the reference's `models` module: `from models import Generator, Discriminator` (trainer.py:11,
tester.py:11)."""
from uegan_b200.models import *  # noqa: F401,F403
from uegan_b200.models import Generator, Discriminator  # noqa: F401
