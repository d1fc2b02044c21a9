"""Stand-in for the top-level jax namespace (forward-only, single device)."""
from . import numpy  # noqa: F401
from . import random  # noqa: F401
from . import lax  # noqa: F401
from . import nn  # noqa: F401


def jit(fn=None, **kwargs):
  if fn is None:
    return lambda f: f
  return fn


def _unsupported(name):
  def f(*a, **k):
    raise NotImplementedError(f'jaxshim: jax.{name} is not implemented')
  return f


def vmap(fn, in_axes=0, out_axes=0):
  """Loop stand-in for jax.vmap over axis 0 of every argument (training.py:180 maps
  compute_elastic_loss over (rays, samples)); tuple outputs are stacked per element."""
  import numpy as _np
  if in_axes != 0 or out_axes != 0:
    raise NotImplementedError('jaxshim jax.vmap: axis 0 only')

  def mapped(*args):
    outs = [fn(*[a[i] for a in args]) for i in range(args[0].shape[0])]
    if isinstance(outs[0], (tuple, list)):
      return tuple(numpy._narrow(_np.stack([_np.asarray(o[k]) for o in outs])) for k in range(len(outs[0])))
    return numpy._narrow(_np.stack([_np.asarray(o) for o in outs]))
  return mapped


def jacfwd(fn, argnums=0):
  """NUMERICAL stand-in for jax.jacfwd (warping.py:196, 385): central differences of the
  reference's own function in float64 (numpy.x64 mode: no narrowing to float32), step 1e-7
  (round-off ~1e-9; the function is piecewise smooth - a step that straddles a ReLU kink of the
  warp MLP gives an O(step) wrong column, which a small step makes rare).  It pins WHICH function
  is differentiated with respect to WHAT and the (output, input) index order of the result."""
  import numpy as _np
  if argnums != 0:
    raise NotImplementedError('jaxshim jax.jacfwd: argnums=0 only')

  def jac(x, *rest):
    with numpy.x64():
      x = _np.asarray(x, dtype=_np.float64)
      cols = []
      for j in range(x.shape[-1]):
        d = _np.zeros_like(x)
        d[..., j] = 1e-7
        hi = _np.asarray(fn(x + d, *rest), dtype=_np.float64)
        lo = _np.asarray(fn(x - d, *rest), dtype=_np.float64)
        cols.append((hi - lo) / 2e-7)
      return _np.stack(cols, axis=-1)
  return jac


def custom_jvp(fn=None, nondiff_argnums=()):
  """utils.py:34 decorates safe_norm; forward-only here."""
  def wrap(f):
    f.defjvp = lambda g: g
    return f
  return wrap if fn is None else wrap(fn)


def device_get(x):
  return x


def process_index():
  return 0


def process_count():
  return 1


def local_device_count():
  return 1


def tree_map(f, tree, *rest):
  if isinstance(tree, dict):
    return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return type(tree)(tree_map(f, v, *[r[i] for r in rest])
                      for i, v in enumerate(tree))
  return f(tree, *rest)


tree_multimap = tree_map


class tree_util:
  tree_map = staticmethod(tree_map)
  tree_multimap = staticmethod(tree_map)
