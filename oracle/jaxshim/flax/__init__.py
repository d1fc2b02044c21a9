"""Stand-in for flax (only what the nerfies hot path imports)."""
from . import linen  # noqa: F401
from . import optim  # noqa: F401
from . import struct  # noqa: F401
