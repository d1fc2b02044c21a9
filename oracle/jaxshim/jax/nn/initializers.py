"""Stand-in for jax.nn.initializers; `key` is a numpy Generator."""
import numpy as _np


def zeros(key, shape, dtype=_np.float32):
  return _np.zeros(shape, dtype)


def uniform(scale=1e-2):
  def init(key, shape, dtype=_np.float32):
    return (key.random(size=tuple(shape)) * scale).astype(dtype)
  return init


def glorot_uniform():
  def init(key, shape, dtype=_np.float32):
    fan_in, fan_out = shape[-2], shape[-1]
    a = _np.sqrt(6.0 / (fan_in + fan_out))
    return ((key.random(size=tuple(shape)) * 2 - 1) * a).astype(dtype)
  return init


xavier_uniform = glorot_uniform
