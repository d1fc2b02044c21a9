"""Stand-in for jax.random: keys are numpy Generators (draws are NOT threefry)."""
import numpy as _np


def PRNGKey(seed):
  return _np.random.default_rng(int(seed))


def split(key, num=2):
  seeds = key.integers(0, 2**31 - 1, size=num)
  return [_np.random.default_rng(int(s)) for s in seeds]


def uniform(key, shape, dtype=_np.float32, minval=0.0, maxval=1.0):
  u = key.random(size=tuple(shape)).astype(_np.float32)
  return (u * (maxval - minval) + minval).astype(dtype)


def normal(key, shape, dtype=_np.float32):
  return key.standard_normal(size=tuple(shape)).astype(dtype)
