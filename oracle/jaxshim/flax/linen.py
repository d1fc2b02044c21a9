"""A miniature flax.linen: Module / compact / setup / Dense / Embed / vmap.

Implements the naming rules the reference relies on:
  * submodules assigned in `setup` are named after the attribute; dict-valued
    attributes name their members `<attr>_<key>`;
  * submodules constructed inside an `@nn.compact` method take their explicit
    `name=` or the auto-name `<ClassName>_<i>` in construction order;
  * `nn.vmap(cls, variable_axes={'params': None}, ...)` shares the wrapped
    module's parameters (same scope, same names) and maps `__call__` over axis 0
    of the arguments whose in_axes entry is 0 - here with a Python loop.
Parameters live in one nested dict {'params': {...}} owned by the root module.
"""
import dataclasses
import functools
from typing import Any, Optional

import numpy as np

from jax.nn import relu, sigmoid, softplus, elu, leaky_relu, tanh  # noqa: F401
from jax.nn import initializers  # noqa: F401

_STACK = []  # modules whose wrapped __call__ is executing.


def compact(fn):
  fn._nn_compact = True
  return fn


def _wrap_call(fn):
  is_compact = getattr(fn, '_nn_compact', False)

  @functools.wraps(fn)
  def wrapped(self, *args, **kwargs):
    self._ensure_setup()
    prev_compact = self.__dict__.get('_in_compact', False)
    prev_counts = self.__dict__.get('_autoname', None)
    object.__setattr__(self, '_in_compact', is_compact)
    object.__setattr__(self, '_autoname', {})
    _STACK.append(self)
    try:
      return fn(self, *args, **kwargs)
    finally:
      _STACK.pop()
      object.__setattr__(self, '_in_compact', prev_compact)
      object.__setattr__(self, '_autoname', prev_counts)
  wrapped._nn_wrapped = True
  return wrapped


class Module:
  name: Optional[str] = None
  parent: Any = None

  def __init_subclass__(cls, **kwargs):
    super().__init_subclass__(**kwargs)
    ann = dict(cls.__dict__.get('__annotations__', {}))
    # name/parent go last, like flax.
    ann.pop('name', None)
    ann.pop('parent', None)
    ann['name'] = Optional[str]
    ann['parent'] = Any
    cls.__annotations__ = ann
    cls.name = None
    cls.parent = None
    call = cls.__dict__.get('__call__')
    if call is not None and not getattr(call, '_nn_wrapped', False):
      cls.__call__ = _wrap_call(call)
    dataclasses.dataclass(cls, eq=False, repr=False)

  # -- construction ---------------------------------------------------------
  def __post_init__(self):
    d = self.__dict__
    d['_setup_done'] = False
    d['_in_setup'] = False
    d['_in_compact'] = False
    d['_autoname'] = None
    d['_variables'] = None
    d['_rngs'] = None
    d['_init_mode'] = False
    if _STACK:
      top = _STACK[-1]
      if top.__dict__.get('_in_compact') and not top.__dict__.get('_in_setup'):
        d['parent'] = top
        if self.name is None:
          counts = top.__dict__['_autoname']
          base = type(self).__dict__.get('_autoname_base', type(self).__name__)
          i = counts.get(base, 0)
          counts[base] = i + 1
          d['name'] = f'{base}_{i}'

  def __setattr__(self, key, value):
    if self.__dict__.get('_in_setup'):
      if isinstance(value, Module):
        value.__dict__['parent'] = self
        value.__dict__['name'] = key
      elif isinstance(value, dict):
        for k, v in value.items():
          if isinstance(v, Module):
            v.__dict__['parent'] = self
            v.__dict__['name'] = f'{key}_{k}'
    object.__setattr__(self, key, value)

  def __getattr__(self, key):
    # only reached when normal lookup fails: attributes defined by setup().
    if key.startswith('__') or '_setup_done' not in self.__dict__:
      raise AttributeError(key)
    if not self.__dict__['_setup_done']:
      self._ensure_setup()
      if key in self.__dict__:
        return self.__dict__[key]
    raise AttributeError(f'{type(self).__name__} has no attribute {key!r}')

  def _ensure_setup(self):
    d = self.__dict__
    if d['_setup_done'] or d['_in_setup']:
      return
    d['_in_setup'] = True
    try:
      self.setup()
    finally:
      d['_in_setup'] = False
    d['_setup_done'] = True

  def setup(self):
    pass

  # -- scope ------------------------------------------------------------------
  def _root(self):
    m = self
    while m.__dict__.get('parent') is not None:
      m = m.__dict__['parent']
    return m

  @property
  def path(self):
    names = []
    m = self
    while m.__dict__.get('parent') is not None:
      names.append(m.__dict__['name'])
      m = m.__dict__['parent']
    return tuple(reversed(names))

  def param(self, name, init_fn, *init_args):
    root = self._root()
    tree = root.__dict__['_variables']['params']
    path = self.path
    if root.__dict__['_init_mode']:
      for p in path:
        tree = tree.setdefault(p, {})
      if name not in tree:
        tree[name] = init_fn(root.__dict__['_rngs']['params'], *init_args)
      return tree[name]
    for p in path:
      if p not in tree:
        raise KeyError(f'missing parameter scope {"/".join(path)} ({p!r})')
      tree = tree[p]
    value = tree[name]
    shape = tuple(init_args[0]) if init_args else None
    if shape is not None and tuple(value.shape) != shape:
      raise ValueError(f'param {"/".join(path)}/{name}: shape {value.shape} '
                       f'!= expected {shape}')
    return value

  def make_rng(self, name):
    rngs = self._root().__dict__['_rngs'] or {}
    if name not in rngs:
      raise KeyError(f'rng {name!r} was not provided')
    return rngs[name]

  # -- entry points -----------------------------------------------------------
  def clone(self):
    kwargs = {f.name: getattr(self, f.name)
              for f in dataclasses.fields(self) if f.name != 'parent'}
    return type(self)(**kwargs)

  def _bind(self, variables, rngs, init_mode):
    saved = list(_STACK)
    _STACK.clear()  # a top-level module must not adopt a caller as parent.
    try:
      m = self.clone()
    finally:
      _STACK.extend(saved)
    m.__dict__['parent'] = None
    m.__dict__['_variables'] = variables
    m.__dict__['_rngs'] = rngs
    m.__dict__['_init_mode'] = init_mode
    return m

  def apply(self, variables, *args, rngs=None, mutable=False, method=None,
            **kwargs):
    m = self._bind({'params': variables['params']}, rngs or {}, False)
    fn = getattr(m, method.__name__) if method is not None else m
    return fn(*args, **kwargs)

  def init(self, rngs, *args, **kwargs):
    if not isinstance(rngs, dict):
      rngs = {'params': rngs}
    variables = {'params': {}}
    m = self._bind(variables, rngs, True)
    m(*args, **kwargs)
    return variables


class Dense(Module):
  features: int
  use_bias: bool = True
  kernel_init: Any = initializers.glorot_uniform()  # flax default is lecun_normal; the reference always overrides or (logit with output_init=None) see below.
  bias_init: Any = initializers.zeros

  def __call__(self, inputs):
    kernel_init = self.kernel_init
    if kernel_init is None:
      raise ValueError('Dense: kernel_init=None')
    kernel = self.param('kernel', kernel_init,
                        (inputs.shape[-1], self.features))
    y = np.matmul(inputs, kernel)
    if self.use_bias:
      bias = self.param('bias', self.bias_init, (self.features,))
      y = y + bias
    return y


class Embed(Module):
  num_embeddings: int
  features: int
  embedding_init: Any = initializers.uniform(1.0)

  def setup(self):
    self.embedding = self.param('embedding', self.embedding_init,
                                (self.num_embeddings, self.features))

  def __call__(self, inputs):
    if not np.issubdtype(np.asarray(inputs).dtype, np.integer):
      raise ValueError('Input type must be an integer or unsigned integer.')
    return self.embedding[np.asarray(inputs).astype(np.int64)]


def _tree_stack(outs):
  first = outs[0]
  if isinstance(first, dict):
    return {k: _tree_stack([o[k] for o in outs]) for k in first}
  if isinstance(first, (tuple, list)):
    return type(first)(_tree_stack([o[i] for o in outs])
                       for i in range(len(first)))
  return np.stack(outs, axis=0)


def vmap(module_cls, variable_axes=None, split_rngs=None, in_axes=0,
         out_axes=0):
  """Lifted vmap: same parameters/scope, `__call__` mapped over axis 0."""
  if out_axes != 0:
    raise NotImplementedError('jaxshim nn.vmap: out_axes must be 0')
  if variable_axes != {'params': None} or split_rngs != {'params': False}:
    raise NotImplementedError('jaxshim nn.vmap: params must be shared')
  inner_call = module_cls.__call__

  def __call__(self, *args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (
        in_axes,) * len(args)
    if len(axes) != len(args):
      raise ValueError('nn.vmap: in_axes / argument count mismatch')
    n = None
    for a, ax in zip(args, axes):
      if ax == 0:
        n = a.shape[0]
        break
    outs = []
    for i in range(n):
      sub = [a[i] if ax == 0 else a for a, ax in zip(args, axes)]
      outs.append(inner_call(self, *sub))
    return _tree_stack(outs)

  base = module_cls.__dict__.get('_autoname_base', module_cls.__name__)
  return type('Vmap' + module_cls.__name__, (module_cls,),
              {'__call__': __call__, '_autoname_base': base,
               '__annotations__': {}})
