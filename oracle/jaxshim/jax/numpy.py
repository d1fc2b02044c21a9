"""numpy-backed stand-in for jax.numpy with JAX's float32-by-default rule.

Every function is numpy's; any float64 (or int64) array it returns is narrowed
to float32 (int32), which is what JAX does with x64 disabled.
"""
import numpy as _np

pi = _np.pi
inf = _np.inf
newaxis = None
ndarray = _np.ndarray
float32 = _np.float32
uint32 = _np.uint32
int32 = _np.int32
bool_ = _np.bool_
shape = _np.shape
linalg = None  # set below


class JArray(_np.ndarray):
  """ndarray whose augmented assignments rebind instead of mutating, like an
  immutable jnp array (`weights += eps` in model_utils.py:156 must not alter
  the caller's coarse weights)."""

  def __iadd__(self, o):
    return _np.add(self, o)

  def __isub__(self, o):
    return _np.subtract(self, o)

  def __imul__(self, o):
    return _np.multiply(self, o)

  def __itruediv__(self, o):
    return _np.true_divide(self, o)


_X64 = [False]


class x64:
  """Context manager: float64 results are kept (jax.config jax_enable_x64); used by the
  numerical jacfwd stand-in only."""

  def __enter__(self):
    self.prev = _X64[0]
    _X64[0] = True

  def __exit__(self, *exc):
    _X64[0] = self.prev


def _narrow(x):
  if _X64[0]:
    if isinstance(x, _np.ndarray) and not isinstance(x, JArray):
      x = x.view(JArray)
    return x
  if isinstance(x, _np.ndarray) or isinstance(x, _np.generic):
    if x.dtype == _np.float64:
      x = x.astype(_np.float32)
    elif x.dtype == _np.int64:
      x = x.astype(_np.int32)
    if isinstance(x, _np.ndarray) and not isinstance(x, JArray):
      x = x.view(JArray)
    return x
  if isinstance(x, (tuple, list)):
    return type(x)(_narrow(v) for v in x)
  return x


def _prep(a):
  """Python scalars / lists entering an array op become float32 arrays."""
  if isinstance(a, (list, tuple)) and a and not isinstance(
      a[0], _np.ndarray) and not isinstance(a[0], (list, tuple)):
    return _narrow(_np.asarray(a))
  return a


def _wrap(fn):
  def wrapped(*args, **kwargs):
    return _narrow(fn(*args, **kwargs))
  wrapped.__name__ = getattr(fn, '__name__', 'fn')
  return wrapped


def array(x, dtype=None):
  a = _np.asarray(x, dtype=dtype)
  return _narrow(a) if dtype is None else a


asarray = array


def broadcast_to(x, shape):
  return _narrow(_np.broadcast_to(array(x), tuple(shape)))


def concatenate(xs, axis=0):
  return _narrow(_np.concatenate([array(x) for x in xs], axis=axis))


def stack(xs, axis=0):
  return _narrow(_np.stack([array(x) for x in xs], axis=axis))


def zeros(shape, dtype=float32):
  return _narrow(_np.zeros(
      tuple(shape) if not isinstance(shape, int) else shape, dtype))


def ones(shape, dtype=float32):
  return _narrow(_np.ones(
      tuple(shape) if not isinstance(shape, int) else shape, dtype))


def reshape(x, shape):
  return _narrow(_np.reshape(x, shape))


def split(x, indices_or_sections, axis=0):
  if isinstance(indices_or_sections, tuple):
    indices_or_sections = list(indices_or_sections)
  return _np.split(x, indices_or_sections, axis=axis)


def where(c, a, b):
  a_arr = isinstance(a, _np.ndarray)
  b_arr = isinstance(b, _np.ndarray)
  # weak-typed python scalars take the other operand's dtype.
  if a_arr and not b_arr:
    b = _np.asarray(b, dtype=a.dtype)
  elif b_arr and not a_arr:
    a = _np.asarray(a, dtype=b.dtype)
  return _narrow(_np.where(c, a, b))


def clip(x, a_min, a_max):
  x = array(x)
  return _narrow(_np.clip(x, _np.asarray(a_min, x.dtype),
                          _np.asarray(a_max, x.dtype)))


def block(x):
  return _narrow(_np.block(x))


class _Linalg:
  @staticmethod
  def svd(x, compute_uv=True, full_matrices=True):
    r = _np.linalg.svd(x, compute_uv=compute_uv, full_matrices=full_matrices)
    return tuple(_narrow(v) for v in r) if compute_uv else _narrow(r)

  @staticmethod
  def det(x):
    return _narrow(_np.linalg.det(x))

  @staticmethod
  def norm(x, axis=None, keepdims=False):
    # jnp.linalg.norm: sqrt(sum(x*x)) in the array's dtype.
    x = array(x)
    return _narrow(_np.sqrt(_np.sum(x * x, axis=axis, keepdims=keepdims)))


linalg = _Linalg()

for _name in ['linspace', 'eye', 'sin', 'cos', 'exp', 'log', 'sqrt', 'tanh',
              'cumsum', 'cumprod', 'sort', 'squeeze', 'expand_dims', 'tile',
              'sum', 'max', 'min', 'maximum', 'minimum', 'argmax',
              'logical_xor', 'ones_like', 'zeros_like', 'full_like', 'arange',
              'abs', 'logaddexp', 'transpose', 'matmul', 'dot', 'mean',
              'square', 'power', 'isnan', 'isfinite', 'all', 'any', 'take',
              'log10', 'floor', 'diag', 'trace', 'log1p', 'expm1', 'greater_equal',
              'take_along_axis', 'cross', 'argsort', 'sign']:
  globals()[_name] = _wrap(getattr(_np, _name))


class _FInfo:
  def __init__(self, dtype):
    self.eps = _np.float32(_np.finfo(dtype).eps)


def finfo(dtype):
  return _FInfo(dtype)
