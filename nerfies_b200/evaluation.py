"""Mirror of nerfies/evaluation.py: render_image, plus the model_fn factory
that eval.py builds with jax.pmap (eval.py:330-348).

Multi-GPU: one process per GPU (torch.distributed, NCCL).  `device_count` is
the number of shards a chunk is split into, exactly as in the reference
(utils.shard, utils.py:334-338); each rank renders shard `rank` and the shards
are exchanged with one all_gather of 24 B/ray - the reference's
lax.all_gather (eval.py:339).  With device_count == 1 there is no collective.
"""
import math
import time

import torch
import torch.distributed as dist

_TREE_TYPES = (dict,)


def _tree_map(fn, tree):
  if isinstance(tree, dict):
    return {k: _tree_map(fn, v) for k, v in tree.items()}
  return fn(tree)


def shard(xs, device_count):
  """utils.shard (utils.py:334-338)."""
  return _tree_map(lambda x: x.reshape((device_count, -1) + tuple(x.shape[1:])), xs)


def unshard(x, padding=0):
  """utils.unshard (utils.py:346-351)."""
  y = x.reshape((x.shape[0] * x.shape[1],) + tuple(x.shape[2:]))
  return y[:-padding] if padding > 0 else y


def make_model_fn(model, **apply_kwargs):
  """The `_model_fn` + pmap + all_gather of eval.py:330-348 as a plain callable
  with the same signature: model_fn(key_0, key_1, params, rays_dict, warp_extra)
  where every leaf of rays_dict has a leading shard axis (device_count, n, ...).
  Returns every shard's output stacked on axis 0, like lax.all_gather."""

  def model_fn(key_0, key_1, params, rays_dict, warp_extra):
    n_shards = rays_dict['origins'].shape[0]
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    if world > 1 and n_shards != world:
      raise ValueError(f'device_count={n_shards} must equal the world size '
                       f'{world}')
    mine = range(n_shards) if world == 1 else [rank]
    outs = []
    for s in mine:
      rays = _tree_map(lambda x: x[s], rays_dict)
      outs.append(model.apply({'params': params}, rays, warp_extra=warp_extra,
                              rngs={'coarse': key_0, 'fine': key_1},
                              mutable=False, **apply_kwargs))
    if world == 1:
      return {lv: {k: torch.stack([o[lv][k] for o in outs], 0)
                   for k in outs[0][lv]} for lv in outs[0]}
    gathered = {}
    for lv, ret in outs[0].items():
      gathered[lv] = {}
      for k, v in ret.items():
        v = v.contiguous()
        buf = torch.empty((world * v.shape[0],) + tuple(v.shape[1:]),
                          device=v.device, dtype=v.dtype)
        dist.all_gather_into_tensor(buf, v)
        gathered[lv][k] = buf.reshape((world,) + tuple(v.shape))
    return gathered

  return model_fn


def render_image(state, rays_dict, model_fn, device_count, rng, chunk=8192,
                 default_ret_key=None):
  """Render all the pixels of an image (evaluation.py:28-101): same arguments,
  chunking, edge padding to a multiple of device_count, and output pytree
  {rgb (h,w,3), depth, med_depth, acc (h,w)}."""
  h, w = rays_dict['origins'].shape[:2]
  rays_dict = _tree_map(lambda x: torch.as_tensor(x).reshape((h * w, -1)),
                        rays_dict)
  num_rays = h * w
  key_0 = key_1 = rng  # keys are unused on the deterministic eval path.
  ret_maps = []
  start_time = time.time()
  num_batches = int(math.ceil(num_rays / chunk))
  for batch_idx in range(num_batches):
    ray_idx = batch_idx * chunk
    chunk_rays = _tree_map(lambda x: x[ray_idx:ray_idx + chunk], rays_dict)
    num_chunk_rays = chunk_rays['origins'].shape[0]
    remainder = num_chunk_rays % device_count
    if remainder != 0:
      padding = device_count - remainder
      # jnp.pad(..., mode='edge'): repeat the last ray.
      chunk_rays = _tree_map(
          lambda x: torch.cat([x, x[-1:].expand(padding, *x.shape[1:])], 0),
          chunk_rays)
    else:
      padding = 0
    chunk_rays = shard(chunk_rays, device_count)
    model_out = model_fn(key_0, key_1, state.optimizer.target['model'],
                         chunk_rays, state.warp_extra)
    if not default_ret_key:
      ret_key = 'fine' if 'fine' in model_out else 'coarse'
    else:
      ret_key = default_ret_key
    ret_map = {k: unshard(v, padding) for k, v in model_out[ret_key].items()}
    ret_maps.append(ret_map)
  out = {}
  for key in ret_maps[0]:
    value = torch.cat([m[key] for m in ret_maps], dim=0)
    out[key] = value.reshape((h, w) + tuple(value.shape[1:]))
  render_image.last_seconds = time.time() - start_time
  return out
