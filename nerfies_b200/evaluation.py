"""Mirror of nerfies/evaluation.py: render_image, plus the model_fn factory
that eval.py builds with jax.pmap (eval.py:330-348), plus a whole-frame renderer.

Multi-GPU: one process per GPU (torch.distributed, NCCL).  `device_count` is
the number of shards a chunk is split into, exactly as in the reference
(utils.shard, utils.py:334-338); each rank renders shard `rank` and the shards
are exchanged with ONE all_gather per chunk of the packed 24 B/ray/level result
(the reference's lax.all_gather, eval.py:339, moves the same values as 4 arrays
per level).  With device_count == 1 there is no collective.

`render_frame` is the B200-native form of the same job (eval.py:330-353 +
datasets/core.py:50-75): every rank generates the rays of its contiguous slab of
the frame on the GPU (camera_rays_kernel), renders it in large launches, and the
frame is assembled with a single all_gather of 24 B/ray at the end - instead of
h*w/chunk host round trips with a collective each.
"""
import math
import time

import torch
import torch.distributed as dist

_TREE_TYPES = (dict,)
_KEYS = (('rgb', slice(0, 3)), ('depth', 3), ('med_depth', 4), ('acc', 5))


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


def _dist_info():
  distributed = dist.is_available() and dist.is_initialized()
  return (dist.get_world_size() if distributed else 1,
          dist.get_rank() if distributed else 0)


def _unpack(packed):
  """(..., 6) -> {'rgb' (...,3), 'depth', 'med_depth', 'acc' (...)} (views)."""
  return {k: packed[..., i] for k, i in _KEYS}


def make_model_fn(model, **apply_kwargs):
  """The `_model_fn` + pmap + all_gather of eval.py:330-348 as a plain callable
  with the same signature: model_fn(key_0, key_1, params, rays_dict, warp_extra)
  where every leaf of rays_dict has a leading shard axis (device_count, n, ...).
  Returns every shard's output stacked on axis 0, like lax.all_gather."""
  simple = not apply_kwargs      # extra outputs (weights, points) take the per-key path

  def model_fn(key_0, key_1, params, rays_dict, warp_extra):
    n_shards = rays_dict['origins'].shape[0]
    world, rank = _dist_info()
    if world > 1 and n_shards != world:
      raise ValueError(f'device_count={n_shards} must equal the world size '
                       f'{world}')
    mine = range(n_shards) if world == 1 else [rank]
    outs = []
    for s in mine:
      rays = _tree_map(lambda x: x[s], rays_dict)
      outs.append(model.apply({'params': params}, rays, warp_extra=warp_extra,
                              rngs={'coarse': key_0, 'fine': key_1},
                              mutable=False, _packed=simple, **apply_kwargs))
    if simple:
      levels = list(outs[0])
      if world == 1:
        return {lv: _unpack(torch.stack([o[lv] for o in outs], 0)) for lv in levels}
      # one collective per chunk: (levels, n, 6) from every rank
      mine_packed = torch.stack([outs[0][lv] for lv in levels], 0).contiguous()
      buf = torch.empty((world * len(levels),) + tuple(mine_packed.shape[1:]),
                        device=mine_packed.device, dtype=mine_packed.dtype)
      dist.all_gather_into_tensor(buf, mine_packed)
      buf = buf.reshape((world, len(levels)) + tuple(mine_packed.shape[1:]))
      return {lv: _unpack(buf[:, i]) for i, lv in enumerate(levels)}
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


def render_frame(model, params, camera, warp_extra, metadata=None, max_rays=65536,
                 default_ret_key=None, timings=None):
  """One full frame from a Camera, B200-native (eval.py:330-353 without the host
  loop): rank r renders pixels [r * n, (r + 1) * n) of the row-major frame,
  n = ceil(h * w / world); rays come from camera_rays_kernel on the device
  (datasets/core.py:50-75), the slab is rendered in launches of up to `max_rays`
  rays, and ONE all_gather of the packed (n, 6) result assembles the frame on
  every rank (padding rays past the end repeat the last pixel, like the
  reference's edge padding, and are dropped).

  metadata: {'warp': id, 'appearance': id, 'camera': id} scalars applied to every
  ray (eval.py:344-348 renders one camera with one metadata id per frame).
  timings (optional dict) receives {'render_ms', 'gather_ms'} measured with CUDA
  events on the current stream.  Returns {rgb (h,w,3), depth, med_depth, acc}."""
  from nerfies_b200 import camera as camera_lib
  world, rank = _dist_info()
  dev = model.device
  h, w = camera.image_shape
  total = h * w
  per = -(-total // world)
  first = rank * per
  count = max(0, min(per, total - first))
  md = metadata or {}
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timings is not None else None
  if ev:
    ev[0].record()
  packed = torch.empty(per, 6, device=dev)
  levels_key = None
  done = 0
  while done < count:
    n = min(max_rays, count - done)
    rays = camera_lib.camera_to_rays(camera, dev, first_pixel=first + done, count=n)
    rays_dict = {'origins': rays['origins'], 'directions': rays['directions'],
                 'metadata': {k: torch.full((n, 1), int(v), dtype=torch.int32, device=dev)
                              for k, v in md.items()}}
    out = model.apply({'params': params}, rays_dict, warp_extra=warp_extra, _packed=True)
    if levels_key is None:
      levels_key = default_ret_key or ('fine' if 'fine' in out else 'coarse')
    packed[done:done + n] = out[levels_key]
    done += n
  if count < per:
    # edge padding (this rank's slab runs past the frame): repeat the last pixel's result
    fill = packed[count - 1:count] if count > 0 else torch.zeros(1, 6, device=dev)
    packed[count:] = fill
  if ev:
    ev[1].record()
  if world > 1:
    buf = torch.empty(world * per, 6, device=dev)
    dist.all_gather_into_tensor(buf, packed)
  else:
    buf = packed
  if ev:
    ev[2].record()
    ev[2].synchronize()
    timings['render_ms'] = ev[0].elapsed_time(ev[1])
    timings['gather_ms'] = ev[1].elapsed_time(ev[2])
  frame = buf[:total]
  return {k: frame[:, i].reshape((h, w) + ((3,) if k == 'rgb' else ()))
          for k, i in _KEYS}
