"""Host-side logic that needs no GPU: parameter pytrees, render_image chunking
and padding, and the one-process-per-GPU sharding of make_model_fn (gloo, world
size 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nerfies_b200 as nb
from nerfies_b200 import evaluation
from nerfies_b200.model_utils import Optimizer, TrainState
from oracle import nerfies_oracle as O
from tests.golden_util import Golden, flatten, model_from_spec


def test_construct_nerf_param_tree_matches_reference_names():
  g = Golden('se3_small')
  cfg = nb.configs.ModelConfig(
      use_warp=True, warp_field_type='se3', use_appearance_metadata=True,
      use_camera_metadata=True, nerf_trunk_width=64, nerf_rgb_branch_width=32,
      num_nerf_point_freqs=8, warp_kwargs={'trunk_width': 32})
  model, params = nb.construct_nerf(0, cfg, 16, range(5), range(2), range(7),
                                    0.02, 0.83, device='cpu')
  ours = {k: tuple(v.shape) for k, v in flatten(params).items()}
  ref = {k: tuple(v.shape) for k, v in flatten(g.params).items()}
  assert ours == ref
  flat = flatten(params)
  heads = flat['warp_field/branches_w/logit/kernel']
  assert float(heads.min()) >= 0 and float(heads.max()) < 1e-4
  emb = flat['warp_field/metadata_encoder/embed/embedding']
  assert float(emb.min()) >= 0 and float(emb.max()) < 0.05
  k = flat['nerf_mlps_coarse/MLP_0/hidden_4/kernel']
  assert float(k.abs().max()) <= np.sqrt(6.0 / sum(k.shape)) + 1e-7
  assert float(flat['nerf_mlps_fine/MLP_0/hidden_0/bias'].abs().max()) == 0
  assert model.num_warp_embeddings == 7 and model.warp_trunk_width == 32


def test_apply_without_gpu_raises():
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  g = Golden('se3_small')
  model = model_from_spec(g.spec_dict, device='cpu')
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    model.apply({'params': g.params}, g.rays, warp_extra={'alpha': 0.0})


class _FakeModel:
  """Stands in for NerfModel on the CPU: a closed-form 'render'."""

  def apply(self, variables, rays, warp_extra=None, rngs=None, mutable=False, _packed=False):
    o, d = rays['origins'], rays['directions']
    scale = variables['params']['scale']
    rgb = (o * 2 + d) * scale + warp_extra['alpha']
    pack = lambda rgb, acc: torch.cat([rgb, o[:, 1:2], d[:, 2:3], acc[:, None]], -1)
    packed = {'coarse': pack(rgb * 0.5, o[:, 0]), 'fine': pack(rgb, o[:, 0] + d[:, 1])}
    if _packed:          # (n,6) per level: rgb3, depth, med_depth, acc - what NerfModel.apply hands over
      return packed
    return {lv: {'rgb': p[:, :3], 'depth': p[:, 3], 'med_depth': p[:, 4], 'acc': p[:, 5]}
            for lv, p in packed.items()}


def _frame(h, w):
  g = torch.Generator().manual_seed(0)
  return {'origins': torch.rand(h, w, 3, generator=g),
          'directions': torch.rand(h, w, 3, generator=g),
          'metadata': {'warp': torch.randint(0, 5, (h, w, 1), generator=g)}}


def _state(scale=3.0, alpha=0.25):
  return TrainState(Optimizer({'model': {'scale': scale}}), warp_alpha=alpha)


def test_render_image_chunks_and_pads_like_the_reference():
  # evaluation.py:52-99: 7x5 = 35 rays, chunk 8, device_count 3 -> every chunk
  # is edge-padded to a multiple of 3 and the padding is dropped again.
  rays = _frame(7, 5)
  fn = evaluation.make_model_fn(_FakeModel())
  out = evaluation.render_image(_state(), rays, fn, device_count=3, rng=0,
                                chunk=8)
  exp = (rays['origins'] * 2 + rays['directions']) * 3.0 + 0.25
  assert out['rgb'].shape == (7, 5, 3) and out['acc'].shape == (7, 5)
  assert torch.allclose(out['rgb'], exp)
  coarse = evaluation.render_image(_state(), rays, fn, 3, 0, chunk=8,
                                   default_ret_key='coarse')
  assert torch.allclose(coarse['rgb'], exp * 0.5)


def test_shard_unshard_roundtrip():
  x = torch.arange(24.).reshape(12, 2)
  s = evaluation.shard({'a': x}, 4)['a']
  assert s.shape == (4, 3, 2)
  assert torch.equal(evaluation.unshard(s), x)
  assert torch.equal(evaluation.unshard(s, padding=2), x[:-2])


def _worker(rank, world, port, tmp):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                    RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    rays = _frame(6, 5)
    fn = evaluation.make_model_fn(_FakeModel())
    out = evaluation.render_image(_state(), rays, fn, device_count=world,
                                  rng=0, chunk=7)
    torch.save(out, os.path.join(tmp, f'out{rank}.pt'))
  finally:
    dist.destroy_process_group()


def test_two_rank_render_equals_single_process(tmp_path):
  """N>1 path: each rank renders its shard, one all_gather assembles the chunk
  (eval.py:339); every rank ends with the full frame, equal to the 1-rank one."""
  port = 29500 + os.getpid() % 2000
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  rays = _frame(6, 5)
  single = evaluation.render_image(
      _state(), rays, evaluation.make_model_fn(_FakeModel()), 1, 0, chunk=7)
  for r in range(2):
    out = torch.load(os.path.join(str(tmp_path), f'out{r}.pt'))
    assert set(out) == {'rgb', 'depth', 'med_depth', 'acc'}
    for k in single:
      assert torch.equal(out[k], single[k]), (r, k)


def test_oracle_and_product_agree_on_condition_widths():
  for kw in [dict(use_viewdirs=True), dict(use_viewdirs=False),
             dict(use_appearance_metadata=True, use_alpha_condition=True),
             dict(use_appearance_metadata=True, use_camera_metadata=True)]:
    spec = O.OracleSpec(**kw)
    p = O.init_params(spec, 0)
    from tests.golden_util import spec_to_dict
    model = model_from_spec(spec_to_dict(spec), device='cpu')
    ours = nb.models.init_params(model, 0)
    assert {k: tuple(v.shape) for k, v in flatten(ours).items()} == {
        k: tuple(v.shape) for k, v in flatten(p).items()}
