"""nerfies_b200: B200-native (sm_100a) render hot path of google/nerfies.

Public surface (mirrors the reference's):
  nerfies_b200.configs     ModelConfig / TrainConfig / EvalConfig + gin subset
  nerfies_b200.models      construct_nerf, NerfModel.apply, WarpField.apply
  nerfies_b200.evaluation  render_image
  nerfies_b200.model_utils TrainState
  nerfies_b200.training    train_step (value_and_grad + gradient all-reduce + Adam)
The arithmetic lives in libnerfies_b200.so (include/nerfies_b200.h); there is no
CPU or PyTorch fallback.
"""
from nerfies_b200 import configs  # noqa: F401
from nerfies_b200 import models  # noqa: F401
from nerfies_b200 import model_utils  # noqa: F401
from nerfies_b200 import evaluation  # noqa: F401
from nerfies_b200 import camera  # noqa: F401
from nerfies_b200 import checkpoints  # noqa: F401
from nerfies_b200 import training  # noqa: F401
from nerfies_b200.models import construct_nerf, NerfModel  # noqa: F401

__version__ = '0.1'
