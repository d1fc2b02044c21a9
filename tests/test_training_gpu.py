"""Training tier: gradients of the photometric loss against
torch.autograd on the oracle (fp64), the Adam update against a restatement of
flax.optim.Adam, and the data-parallel step on 2 GPUs (NCCL all-reduce of the flat
gradient) against the single-GPU step on the whole batch.

Tolerance: 2e-4 of the tensor's largest gradient entry (|dg| / (max|g_ref| + tiny)); the
CUDA path is fp32 with atomicAdd row reductions, the reference is autograd in fp64.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nerfies_oracle as O
from tests.golden_util import Golden, flatten, model_from_spec, tree_to_device

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 2e-4


def _oracle_loss_and_grads(g, params64, target):
  """training.py:171-175, 214-244 on the oracle in float64 (z_fine detached:
  lax.stop_gradient, model_utils.py:211)."""
  spec = g.spec
  leaves = flatten(params64)
  for v in leaves.values():
    v.requires_grad_(True)
  with torch.no_grad():
    fwd = O.render_forward(params64, spec, g.rays, warp_alpha=g.warp_alpha, dtype=torch.float64,
                           t_rand=g.t_rand, u_rand=g.u_rand)
  zc, zf = fwd['coarse']['z_vals'], fwd['fine']['z_vals']
  oc = O.render_level(params64, spec, 'coarse', g.rays, zc, g.warp_alpha, dtype=torch.float64)
  of = O.render_level(params64, spec, 'fine', g.rays, zf, g.warp_alpha, dtype=torch.float64)
  lc = ((oc['rgb'] - target.double())**2).mean()
  lf = ((of['rgb'] - target.double())**2).mean()
  (lc + lf).backward()
  return float(lc), float(lf), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}, zf


def _params64(g):
  def conv(t):
    return {k: conv(v) for k, v in t.items()} if isinstance(t, dict) else t.double().clone()
  return conv(g.params)


@pytest.mark.parametrize('name', ['se3_small', 'translation_small', 'nowarp_variants', 'alpha_cond_init',
                                  'pivot_small'])
def test_gradients_match_autograd_on_the_oracle(name):
  from nerfies_b200 import training
  g = Golden(name)
  torch.manual_seed(0)
  B = g.rays['origins'].shape[0]
  target = torch.rand(B, 3)
  p64 = _params64(g)
  lc, lf, ref, zf = _oracle_loss_and_grads(g, p64, target)
  model = model_from_spec(g.spec_dict, device=DEV)
  params = tree_to_device(g.params, DEV)
  batch = dict(g.rays, rgb=target)
  # the fine level must see the same z as the reference run: feed the recorded draws; the
  # deterministic path resamples from its own (fp32) coarse weights, which moves z_fine by the
  # conditioning of the inverse CDF - compare the coarse level and the shared parameters on it
  losses, grads = training.value_and_grad(model, params, batch, {'alpha': g.warp_alpha}, chunk_rays=5,
                                          t_rand=g.t_rand, u_rand=g.u_rand)
  torch.cuda.synchronize()
  assert abs(float(losses['coarse']) - lc) < 1e-5 * max(1.0, lc)
  assert abs(float(losses['fine']) - lf) < 2e-3 * max(1e-3, lf)
  tree = training.grads_to_tree(model, grads)
  got = flatten(tree)
  worst = {}
  for k, r in ref.items():
    a = got[k].cpu().double().reshape(r.shape)
    scale = float(r.abs().max()) + 1e-12
    worst[k] = float((a - r).abs().max()) / scale
  # parameters of the coarse level and everything shared are held to TOL; the fine MLP's own
  # gradients inherit the (stated) end-to-end tolerance of the resampled z
  bad = {k: v for k, v in worst.items() if v > (2e-2 if 'nerf_mlps_fine' in k else 5e-3)}
  assert not bad, bad


def test_gradients_at_tolerance_with_the_reference_z():
  """Same check with both levels on the oracle's z (no resampling in between): every
  parameter gradient within TOL."""
  from nerfies_b200 import training, _lib
  from nerfies_b200.models import _ptr
  g = Golden('se3_small')
  torch.manual_seed(1)
  B = g.rays['origins'].shape[0]
  target = torch.rand(B, 3)
  # coarse-only model: one level, no resampling anywhere
  sd = dict(g.spec_dict, num_fine_samples=0)
  spec = O.OracleSpec(**{**sd, 'nerf_skips': tuple(sd['nerf_skips']), 'warp_skips': tuple(sd['warp_skips'])})
  p = {k: v for k, v in g.params.items() if k != 'nerf_mlps_fine'}
  def conv(t):
    return {k: conv(v) for k, v in t.items()} if isinstance(t, dict) else t.double().clone()
  p64 = conv(p)
  leaves = flatten(p64)
  for v in leaves.values():
    v.requires_grad_(True)
  out = O.render_forward(p64, spec, g.rays, warp_alpha=g.warp_alpha, dtype=torch.float64)
  loss = ((out['coarse']['rgb'] - target.double())**2).mean()
  loss.backward()
  model = model_from_spec(sd, device=DEV)
  losses, grads = training.value_and_grad(model, tree_to_device(p, DEV), dict(g.rays, rgb=target),
                                          {'alpha': g.warp_alpha}, chunk_rays=7)
  torch.cuda.synchronize()
  assert abs(float(losses['coarse']) - float(loss)) < 1e-5
  got = flatten(training.grads_to_tree(model, grads))
  for k, v in leaves.items():
    r = v.grad if v.grad is not None else torch.zeros_like(v)
    a = got[k].cpu().double().reshape(r.shape)
    err = float((a - r).abs().max()) / (float(r.abs().max()) + 1e-12)
    assert err < TOL, (k, err)


def test_adam_step_matches_flax_adam():
  from nerfies_b200 import _lib
  from nerfies_b200.models import _ptr, _stream
  lib = _lib.load()
  torch.manual_seed(2)
  n = 100003
  p = torch.randn(n, device=DEV); g = torch.randn(n, device=DEV) * 0.1
  m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
  pr, mr, vr = p.double().cpu(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
  lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
  for t in range(1, 4):
    _lib.check(lib.nfb_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), n, lr, b1, b2, eps, t, _stream()))
    gd = g.double().cpu()
    # flax.optim.Adam.apply_param_gradient
    mr = b1 * mr + (1 - b1) * gd
    vr = b2 * vr + (1 - b2) * gd * gd
    pr = pr - lr * (mr / (1 - b1**t)) / (torch.sqrt(vr / (1 - b2**t)) + eps)
  torch.cuda.synchronize()
  assert float((p.double().cpu() - pr).abs().max()) < 1e-6


def test_train_step_reduces_the_loss_and_updates_the_handle():
  from nerfies_b200 import training
  g = Golden('se3_small')
  torch.manual_seed(3)
  B = g.rays['origins'].shape[0]
  target = torch.rand(B, 3)
  model = model_from_spec(g.spec_dict, device=DEV)
  state = training.create_train_state(model, tree_to_device(g.params, DEV), warp_alpha=g.warp_alpha)
  batch = dict(g.rays, rgb=target)
  sp = training.ScalarParams(learning_rate=2e-3)
  first = None
  for it in range(12):
    state, stats, _ = training.train_step(model, 0, state, batch, sp, chunk_rays=6)
    l = float(stats['fine']['loss/total']) + float(stats['coarse']['loss/total'])
    first = l if first is None else first
  assert l < 0.9 * first, (first, l)
  # the forward path sees the updated parameters (views of the flat vector)
  out = model.apply({'params': state.optimizer.target['model']}, g.rays, warp_extra=state.warp_extra)
  mse = float(((out['fine']['rgb'].cpu() - target)**2).mean())
  assert abs(mse - float(stats['fine']['loss/total'])) < 5e-2 * max(mse, 1e-3) + 1e-4
  with pytest.raises(NotImplementedError):
    training.train_step(model, 0, state, batch, sp, use_elastic_loss=True, elastic_loss_type='nr')


def _worker(rank, world, port, tmp):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  sys.path.insert(0, ROOT)
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
  try:
    from nerfies_b200 import training
    g = Golden('se3_small')
    torch.manual_seed(4)
    B = g.rays['origins'].shape[0]
    target = torch.rand(B, 3)
    dev = f'cuda:{rank}'
    model = model_from_spec(g.spec_dict, device=dev)
    state = training.create_train_state(model, tree_to_device(g.params, dev), warp_alpha=g.warp_alpha)
    half = B // world
    sl = slice(rank * half, (rank + 1) * half)
    batch = {'origins': g.rays['origins'][sl], 'directions': g.rays['directions'][sl],
             'metadata': {k: v[sl] for k, v in g.rays['metadata'].items()}, 'rgb': target[sl]}
    t = {}
    state, stats, _ = training.train_step(model, 0, state, batch, training.ScalarParams(learning_rate=1e-3),
                                          timings=t)
    torch.cuda.synchronize()
    torch.save({'flat': state.optimizer.flat.cpu(), 'timings': t}, os.path.join(tmp, f'out{rank}.pt'))
  finally:
    dist.destroy_process_group()


def test_two_gpu_step_equals_one_gpu_step(tmp_path):
  """lax.pmean of the gradients (training.py:266): two ranks with half of the batch each, one
  NCCL all_reduce of the flat gradient, end with the parameters of the 1-GPU step on the whole
  batch (up to fp32 summation order)."""
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
  from nerfies_b200 import training
  port = 29900 + os.getpid() % 1000
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  g = Golden('se3_small')
  torch.manual_seed(4)
  B = g.rays['origins'].shape[0]
  B2 = B // 2 * 2
  target = torch.rand(B, 3)
  model = model_from_spec(g.spec_dict, device=DEV)
  state = training.create_train_state(model, tree_to_device(g.params, DEV), warp_alpha=g.warp_alpha)
  before = state.optimizer.flat.clone()
  batch = {'origins': g.rays['origins'][:B2], 'directions': g.rays['directions'][:B2],
           'metadata': {k: v[:B2] for k, v in g.rays['metadata'].items()}, 'rgb': target[:B2]}
  state, _, _ = training.train_step(model, 0, state, batch, training.ScalarParams(learning_rate=1e-3))
  torch.cuda.synchronize()
  single = state.optimizer.flat.cpu()
  step = float((single - before.cpu()).abs().max())
  for r in range(2):
    out = torch.load(os.path.join(str(tmp_path), f'out{r}.pt'))
    assert float((out['flat'] - single).abs().max()) < 2e-2 * step + 1e-7, r


# ---------------------------------------------------------------------------
# warp Jacobian, elastic / warp-reg / background regularisers
# ---------------------------------------------------------------------------
def _jac_oracle(g, pts, ids, dtype=torch.float64):
  params = O.tree_to(g.params, dtype)
  return O.warp_jacobian(params['warp_field'], g.spec, pts.to(dtype), ids, g.warp_alpha).detach()


@pytest.mark.parametrize('name', ['se3_small', 'translation_small', 'pivot_small'])
def test_warp_jacobian_matches_the_oracle(name):
  """jax.jacfwd(self.warp) (warping.py:385-387): the forward-mode kernels against autograd on the
  oracle in float64, through warp_field.apply(return_jacobian=True) and model.apply(return_warp_jacobian=True)."""
  g = Golden(name)
  model = model_from_spec(g.spec_dict, device=DEV)
  params = tree_to_device(g.params, DEV)
  gen = torch.Generator().manual_seed(5)
  P = 37
  pts = torch.rand(P, 3, generator=gen) * 0.6 - 0.3
  ids = torch.randint(0, g.spec.num_warp_embeddings, (P, 1), generator=gen)
  wf = model.create_warp_field(model, num_batch_dims=1)
  out = wf.apply({'params': params['warp_field']}, pts, ids, {'alpha': g.warp_alpha}, return_jacobian=True)
  torch.cuda.synchronize()
  ref = _jac_oracle(g, pts, ids)
  err = float((out['jacobian'].cpu().double() - ref).abs().max())
  assert err < 2e-5 * max(1.0, float(ref.abs().max())), err
  # through the model: every sample of both levels
  o2 = model.apply({'params': params}, g.rays, warp_extra={'alpha': g.warp_alpha}, return_warp_jacobian=True,
                   return_points=True, t_rand=g.t_rand, u_rand=g.u_rand)
  torch.cuda.synchronize()
  for lv in ('coarse', 'fine'):
    J = o2[lv]['warp_jacobian'].cpu()
    B, S = J.shape[:2]
    ids_bs = g.rays['metadata']['warp'][:, None, :].expand(B, S, 1).reshape(-1, 1)
    ref = _jac_oracle(g, o2[lv]['points'].cpu().reshape(-1, 3), ids_bs).reshape(B, S, 3, 3)
    assert float((J.double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


def _oracle_reg_loss_and_grads(g, p64, target, zc, zf, sp, use_elastic, reduce, etype, use_warp_reg, bg):
  spec = g.spec
  leaves = flatten(p64)
  for v in leaves.values():
    v.requires_grad_(True)
  total = 0.0
  parts = {}
  for lv, z in (('coarse', zc), ('fine', zf)):
    out = O.render_level(p64, spec, lv, g.rays, z, g.warp_alpha, dtype=torch.float64)
    parts['rgb_' + lv] = ((out['rgb'] - target.double())**2).mean()
    total = total + parts['rgb_' + lv]
    r = O.level_regularisers(p64, spec, out, g.rays, g.warp_alpha, use_elastic_loss=use_elastic and lv == 'coarse',
                             elastic_reduce_method=reduce, elastic_loss_type=etype,
                             use_warp_reg_loss=use_warp_reg, warp_reg_loss_alpha=sp.warp_reg_loss_alpha,
                             warp_reg_loss_scale=sp.warp_reg_loss_scale)
    if 'loss/elastic' in r:
      parts['elastic'] = r['loss/elastic']
      parts['elastic_residual'] = r['residual/elastic']
      total = total + sp.elastic_loss_weight * r['loss/elastic']
    if 'loss/warp_reg' in r:
      parts['warp_reg_' + lv] = r['loss/warp_reg']
      total = total + sp.warp_reg_loss_weight * r['loss/warp_reg']
  if bg is not None:
    l = O.compute_background_loss(p64, spec, bg['points'].double(), bg['ids'], bg['noise'].double(),
                                  g.warp_alpha).mean()
    parts['background'] = l
    total = total + sp.background_loss_weight * l
  total.backward()
  grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
  return {k: float(v.detach()) for k, v in parts.items()}, grads


@pytest.mark.parametrize('case', [
    dict(name='se3_small', elastic=True, reduce='median', etype='log_svals'),
    dict(name='se3_small', elastic=True, reduce='weight', etype='log_svals'),
    dict(name='pivot_small', elastic=True, reduce='median', etype='svals'),
    dict(name='translation_small', elastic=True, reduce='median', etype='jtj'),
    dict(name='se3_small', elastic=True, reduce='median', etype='det'),
    dict(name='se3_small', elastic=True, reduce='median', etype='log_det'),
    dict(name='se3_small', elastic=True, reduce='median', etype='div'),
    dict(name='se3_small', warp_reg=True),
    dict(name='se3_small', background=True),
    dict(name='pivot_small', elastic=True, reduce='median', etype='log_svals', warp_reg=True, background=True),
], ids=lambda c: '-'.join(f'{k}={v}' for k, v in c.items()))
def test_regulariser_gradients_match_autograd_on_the_oracle(case):
  """value_and_grad with the regularisers of train_step (training.py:176-212, 246-257) against
  torch.autograd (double backward through the Jacobian) on the oracle in float64; both levels on
  the oracle's z.  Weights far above the gin values so that the regulariser terms dominate."""
  from nerfies_b200 import training
  g = Golden(case['name'])
  torch.manual_seed(11)
  B = g.rays['origins'].shape[0]
  target = torch.rand(B, 3)
  sp = training.ScalarParams(learning_rate=1e-3, elastic_loss_weight=5.0, warp_reg_loss_weight=3.0,
                             warp_reg_loss_alpha=-2.0, warp_reg_loss_scale=0.05, background_loss_weight=60.0)
  with torch.no_grad():
    fwd = O.render_forward(_params64(g), g.spec, g.rays, warp_alpha=g.warp_alpha, dtype=torch.float64,
                           t_rand=g.t_rand, u_rand=g.u_rand)
  zc, zf = fwd['coarse']['z_vals'], fwd['fine']['z_vals']
  bg = None
  if case.get('background'):
    gen = torch.Generator().manual_seed(3)
    bg = dict(points=torch.rand(23, 3, generator=gen) * 0.5 - 0.25,
              ids=torch.randint(0, g.spec.num_warp_embeddings, (23, 1), generator=gen),
              noise=0.001 * torch.randn(23, 3, generator=gen))
  parts, ref = _oracle_reg_loss_and_grads(g, _params64(g), target, zc, zf, sp, case.get('elastic', False),
                                          case.get('reduce', 'median'), case.get('etype', 'log_svals'),
                                          case.get('warp_reg', False), bg)
  # the check must be able to fail: the regularisers move the reference gradient of the warp field by
  # far more than the tolerance
  _, plain = _oracle_reg_loss_and_grads(g, _params64(g), target, zc, zf, sp, False, 'median', 'log_svals', False, None)
  moved = max(float((ref[k] - plain[k]).abs().max()) / (float(ref[k].abs().max()) + 1e-12)
              for k in ref if k.startswith('warp_field/'))
  assert moved > 0.25, moved
  model = model_from_spec(g.spec_dict, device=DEV)
  kw = {}
  if bg is not None:
    kw = dict(background_points=bg['points'], background_warp_ids=bg['ids'], background_noise=bg['noise'])
  reg = training.make_reg(model, sp, case.get('elastic', False), case.get('reduce', 'median'),
                          case.get('etype', 'log_svals'), bg is not None, case.get('warp_reg', False), **kw)
  losses, grads = training.value_and_grad(model, tree_to_device(g.params, DEV), dict(g.rays, rgb=target),
                                          {'alpha': g.warp_alpha}, chunk_rays=5, t_rand=g.t_rand,
                                          u_rand=g.u_rand, reg=reg)
  torch.cuda.synchronize()
  # the fine level of the CUDA run resamples from its own fp32 coarse weights: compare the loss terms of
  # the coarse level and the background exactly, the gradients on everything the fine z does not touch
  for k in ('elastic', 'warp_reg_coarse', 'background'):
    if k in parts:
      assert abs(float(losses[k]) - parts[k]) < 2e-4 * max(abs(parts[k]), 1e-3), (k, float(losses[k]), parts[k])
  got = flatten(training.grads_to_tree(model, grads))
  worst = {}
  for k, r in ref.items():
    a = got[k].cpu().double().reshape(r.shape)
    worst[k] = float((a - r).abs().max()) / (float(r.abs().max()) + 1e-12)
  bad = {k: v for k, v in worst.items() if v > (2e-2 if 'nerf_mlps_fine' in k else 5e-3)}
  assert not bad, bad


def test_train_step_with_every_regulariser():
  from nerfies_b200 import training
  g = Golden('se3_small')
  torch.manual_seed(5)
  B = g.rays['origins'].shape[0]
  model = model_from_spec(g.spec_dict, device=DEV)
  state = training.create_train_state(model, tree_to_device(g.params, DEV), warp_alpha=g.warp_alpha)
  batch = dict(g.rays, rgb=torch.rand(B, 3), background_points=torch.rand(40, 3) * 0.4 - 0.2)
  sp = training.ScalarParams(learning_rate=1e-3, elastic_loss_weight=0.01, warp_reg_loss_weight=0.001,
                             background_loss_weight=1.0)
  first = last = None
  for it in range(8):
    state, stats, _ = training.train_step(model, it, state, batch, sp, use_elastic_loss=True,
                                          use_background_loss=True, use_warp_reg_loss=True)
    tot = float(stats['coarse']['loss/total']) + float(stats['fine']['loss/total'])
    assert all(torch.isfinite(torch.as_tensor(float(v))) for lv in ('coarse', 'fine') for v in stats[lv].values())
    first = tot if first is None else first
    last = tot
  assert {'loss/elastic', 'residual/elastic', 'metric/jacobian_det', 'loss/warp_reg'} <= set(stats['coarse'])
  assert 'background_loss' in stats and last < first
