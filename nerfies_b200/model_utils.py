"""The caller-visible pieces of nerfies/model_utils.py that are not kernels."""
import dataclasses
from typing import Any


@dataclasses.dataclass
class Optimizer:
  """Stand-in for flax.optim.Optimizer: callers read `.target` only
  (evaluation.py:86 reads state.optimizer.target['model'])."""
  target: Any


@dataclasses.dataclass
class TrainState:
  """model_utils.TrainState (model_utils.py:25-33)."""
  optimizer: Optimizer
  warp_alpha: float = 0.0
  time_alpha: float = 0.0

  @property
  def warp_extra(self):
    return {'alpha': self.warp_alpha, 'time_alpha': self.time_alpha}
