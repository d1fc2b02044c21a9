"""flax.training stand-in (training.py:24 imports checkpoints; unused by the fixtures)."""
from . import checkpoints  # noqa: F401
