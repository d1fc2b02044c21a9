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


vmap = _unsupported('vmap')
jacfwd = _unsupported('jacfwd')


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
