"""Mirror of the optimisation step of nerfies/training.py.

  state = training.create_train_state(model, params)       # flat fp32 parameter / Adam buffers
  state, stats, rng = training.train_step(model, rng, state, batch, scalar_params)

`train_step` has the reference's signature (training.py:138-147) and does what its
body does (training.py:214-271): value_and_grad of the photometric loss
(mean squared error of the coarse and of the fine rgb against batch['rgb']), the
gradient mean over the devices (lax.pmean -> ONE NCCL all_reduce of the flat gradient
vector), and flax.optim.Adam's update.  Gradients come from
`nfb_train_value_and_grad` (hand-written fp32 kernels, include/nerfies_b200.h), the
update from `nfb_adam_step`; torch is the allocator and the NCCL binding.

Not implemented: the elastic loss (needs the warp Jacobian's backward), the warp
regulariser, and the background loss; asking for them raises NotImplementedError.
"""
import ctypes
import dataclasses
import math
from typing import Any, Dict

import torch
import torch.distributed as dist

from nerfies_b200 import _lib
from nerfies_b200 import model_utils
from nerfies_b200.models import _prep_f32, _prep_ids, _ptr, _stream


@dataclasses.dataclass
class ScalarParams:
  """training.ScalarParams (training.py:35-43)."""
  learning_rate: float
  elastic_loss_weight: float = 0.0
  warp_reg_loss_weight: float = 0.0
  warp_reg_loss_alpha: float = -2.0
  warp_reg_loss_scale: float = 0.001
  background_loss_weight: float = 0.0
  background_noise_std: float = 0.001


class AdamOptimizer(model_utils.Optimizer):
  """flax.optim.Adam(learning_rate) wrapped like flax.optim.Optimizer: `.target` is the
  parameter pytree ({'model': params}); its leaves are views into one flat fp32 vector, so the
  update and the gradient all-reduce are single passes over contiguous memory."""

  def __init__(self, target, flat, specs, beta1=0.9, beta2=0.999, eps=1e-8):
    super().__init__(target)
    self.flat = flat
    self.specs = specs                      # [(name, offset, numel)] in nfb_param_info order
    self.m = torch.zeros_like(flat)
    self.v = torch.zeros_like(flat)
    self.step = 0
    self.beta1, self.beta2, self.eps = beta1, beta2, eps

  def apply_gradient(self, grad_flat, learning_rate):
    """optimizer.apply_gradient(grad, learning_rate=...) (training.py:268-269); in place."""
    self.step += 1
    lib = _lib.load()
    with torch.cuda.device(self.flat.device):
      _lib.check(lib.nfb_adam_step(_ptr(self.flat), _ptr(grad_flat), _ptr(self.m), _ptr(self.v),
                                   self.flat.numel(), float(learning_rate), self.beta1, self.beta2,
                                   self.eps, self.step, _stream()))
    return self


def create_train_state(model, params, warp_alpha=0.0, time_alpha=0.0):
  """Flattens the parameter pytree (Flax names) into one fp32 vector in the library's parameter
  order and rebuilds the pytree as views of it; returns model_utils.TrainState with an
  AdamOptimizer (train.py:219-221: optimizer = flax.optim.Adam(lr).create(params))."""
  hd = model.handle(1)
  specs, off = [], 0
  for name, rows, cols in hd.param_specs:
    specs.append((name, off, rows * cols))
    off += rows * cols
  flat = torch.zeros(off, device=model.device, dtype=torch.float32)
  tree: Dict[str, Any] = {}
  for (name, o, n), (_, rows, cols) in zip(specs, hd.param_specs):
    node = params
    for part in name.split('/'):
      node = node[part]
    src = torch.as_tensor(node).to(device=model.device, dtype=torch.float32)
    flat[o:o + n] = src.reshape(-1)
    view = flat[o:o + n].view(src.shape)
    t = tree
    parts = name.split('/')
    for part in parts[:-1]:
      t = t.setdefault(part, {})
    t[parts[-1]] = view
  opt = AdamOptimizer({'model': tree}, flat, specs)
  return model_utils.TrainState(opt, warp_alpha=warp_alpha, time_alpha=time_alpha)


def value_and_grad(model, params, batch, warp_extra, rngs=None, chunk_rays=256, t_rand=None,
                   u_rand=None, grads=None):
  """(loss dict, flat gradient) of the photometric loss (training.py:171-175, 214-244, 263-264).
  `grads` (flat, zeroed by the caller) may be passed to accumulate into."""
  dev = model.device
  origins = _prep_f32(batch['origins'], dev)
  directions = _prep_f32(batch['directions'], dev)
  B = origins.shape[0]
  viewdirs = _prep_f32(batch['viewdirs'], dev) if 'viewdirs' in batch else None
  md = batch.get('metadata', {})
  warp_id = _prep_ids(md.get('warp'), dev) if model.use_warp else None
  app_id = _prep_ids(md.get('appearance'), dev) if model.use_appearance_metadata else None
  cam_id = _prep_ids(md.get('camera'), dev) if model.use_camera_metadata else None
  target = _prep_f32(batch['rgb'], dev)[..., :3].contiguous()
  if t_rand is None and u_rand is None:
    t_rand, u_rand = model._draws(rngs, B)
  t_rand = None if t_rand is None else _prep_f32(t_rand, dev)
  u_rand = None if u_rand is None else _prep_f32(u_rand, dev)
  hd = model.handle(B)
  hd.set_params(params)
  n = len(hd.param_specs)
  numels = [r * c for _, r, c in hd.param_specs]
  if grads is None:
    grads = torch.zeros(sum(numels), device=dev, dtype=torch.float32)
  ptrs, off = (ctypes.c_void_p * n)(), 0
  for i, k in enumerate(numels):
    ptrs[i] = grads.data_ptr() + 4 * off
    off += k
  loss = torch.zeros(2, device=dev)
  with torch.cuda.device(dev):
    _lib.check(hd.lib.nfb_train_value_and_grad(
        hd.h, B, _ptr(origins), _ptr(directions), _ptr(viewdirs), _ptr(warp_id), _ptr(app_id),
        _ptr(cam_id), float((warp_extra or {}).get('alpha', 0.0)), _ptr(t_rand), _ptr(u_rand), 0,
        _ptr(target), int(chunk_rays), ptrs, (ctypes.c_longlong * n)(*numels), n, _ptr(loss),
        _stream()))
  return {'coarse': loss[0], 'fine': loss[1]}, grads


def grads_to_tree(model, grads):
  """Flat gradient -> pytree with the Flax names (views)."""
  hd = model.handle(1)
  tree, off = {}, 0
  for name, rows, cols in hd.param_specs:
    t = tree
    parts = name.split('/')
    for part in parts[:-1]:
      t = t.setdefault(part, {})
    shape = (cols,) if parts[-1] == 'bias' else (rows, cols)
    t[parts[-1]] = grads[off:off + rows * cols].view(shape)
    off += rows * cols
  return tree


def train_step(model, rng_key, state, batch, scalar_params, use_elastic_loss=False,
               elastic_reduce_method='median', elastic_loss_type='log_svals',
               use_background_loss=False, use_warp_reg_loss=False, chunk_rays=256,
               timings=None):
  """One optimisation step (training.py:138-271).  Returns (new_state, stats, rng_key).

  timings (optional dict) receives 'value_and_grad_ms', 'all_reduce_ms', 'adam_ms' measured
  with CUDA events on the current stream."""
  del elastic_reduce_method, elastic_loss_type
  if use_elastic_loss or use_warp_reg_loss or use_background_loss:
    raise NotImplementedError(
        'elastic / warp-reg / background regularisers (training.py:71-135, 187-212, 246-257) are '
        'not implemented: they need the backward of the warp Jacobian (SURVEY §8f #2)')
  opt = state.optimizer
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
  if ev:
    ev[0].record()
  params = opt.target['model']
  model.invalidate_params()     # the Adam kernel rewrote the flat vector under the views
  losses, grads = value_and_grad(model, params, batch, state.warp_extra,
                                 rngs={'coarse': rng_key, 'fine': rng_key}, chunk_rays=chunk_rays)
  if ev:
    ev[1].record()
  stats = {lv: {'loss/rgb': l, 'loss/total': l, 'metric/psnr': -10.0 * torch.log10(l)}
           for lv, l in losses.items()}
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    # jax.lax.pmean(grad, 'batch') (training.py:266): one collective over the flat vector
    dist.all_reduce(grads)
    grads /= dist.get_world_size()
  if ev:
    ev[2].record()
  opt.apply_gradient(grads, learning_rate=scalar_params.learning_rate)
  if ev:
    ev[3].record()
    ev[3].synchronize()
    timings['value_and_grad_ms'] = ev[0].elapsed_time(ev[1])
    timings['all_reduce_ms'] = ev[1].elapsed_time(ev[2])
    timings['adam_ms'] = ev[2].elapsed_time(ev[3])
  return state, stats, rng_key
