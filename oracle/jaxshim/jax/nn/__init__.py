"""Stand-in for jax.nn: activations + initializers, float32 numpy."""
import numpy as _np

from . import initializers  # noqa: F401


def relu(x):
  return _np.maximum(x, _np.zeros((), x.dtype))


def sigmoid(x):
  # jax.nn.sigmoid = lax.logistic = 1 / (1 + exp(-x)).
  one = _np.ones((), x.dtype)
  return one / (one + _np.exp(-x))


def softplus(x):
  # jax.nn.softplus(x) = jnp.logaddexp(x, 0).
  return _np.logaddexp(x, _np.zeros((), x.dtype))


def elu(x, alpha=1.0):
  return _np.where(x > 0, x, (alpha * _np.expm1(_np.minimum(x, 0))).astype(x.dtype))


def leaky_relu(x, negative_slope=1e-2):
  return _np.where(x >= 0, x, (negative_slope * x).astype(x.dtype))


def tanh(x):
  return _np.tanh(x)
