"""flax.optim was removed from modern Flax; model_utils.py:18 imports it."""


class Optimizer:
  target = None


class Adam:
  def __init__(self, *a, **k):
    pass
