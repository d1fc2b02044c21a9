"""Oracle restatement of the training-tier regularisers against known
answers produced by the reference's own source (oracle/make_golden_reg.py ->
tests/golden/regularisers.npz)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import nerfies_oracle as O
from tests.golden_util import GOLDEN_DIR, Golden

Z = np.load(os.path.join(GOLDEN_DIR, 'regularisers.npz'))


@pytest.mark.parametrize('loss_type', ['log_svals', 'svals', 'jtj', 'div', 'det', 'log_det', 'nr'])
def test_compute_elastic_loss(loss_type):
  """training.py:71-115."""
  J = torch.from_numpy(Z['elastic/J'])
  loss, residual = O.compute_elastic_loss(J, loss_type=loss_type)
  ref_l, ref_r = Z[f'elastic/{loss_type}/loss'], Z[f'elastic/{loss_type}/residual']
  # the residual is a sqrt of a difference that vanishes at the identity: absolute floor
  np.testing.assert_allclose(residual.numpy(), ref_r, rtol=2e-4, atol=2e-6)
  np.testing.assert_allclose(loss.numpy(), ref_l, rtol=4e-4, atol=1e-8)


def test_general_loss_with_squared_residual():
  """utils.py:264-331 for alpha in {-inf, -2, -0.5, 0, 1, 2, +inf}."""
  sq = torch.from_numpy(Z['gl/sq'])
  for a, s, ref in zip(Z['gl/alpha'], Z['gl/scale'], Z['gl/loss']):
    got = O.general_loss_with_squared_residual(sq, float(a), float(s)).numpy()
    ok = np.isfinite(ref)
    # (pow(1 + x/b, a/2) - 1 cancels in fp32 for tiny x: absolute floor ~ scale * 1e-7)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=3e-5, atol=float(s) * 6e-7, err_msg=f'alpha={a} scale={s}')


def test_compute_depth_index():
  """model_utils.py:218-245, including rays that never reach the threshold (index 0)."""
  idx = O.compute_depth_index(torch.from_numpy(Z['depth/weights']))
  assert idx.tolist() == Z['depth/index'].tolist()


@pytest.mark.parametrize('name', ['se3_small', 'translation_small', 'pivot_small'])
def test_warp_jacobian(name):
  """warping.py:385-387: autograd on the oracle (float64) against the reference's own `warp`
  differentiated numerically in float64 by the shim's jacfwd (central differences, step 1e-7)."""
  g = Golden(name)
  p64 = O.tree_to(g.params, torch.float64)
  pts = torch.from_numpy(Z[f'jac/{name}/points']).double()
  ids = torch.from_numpy(Z[f'jac/{name}/ids'].astype(np.int64))
  J = O.warp_jacobian(p64['warp_field'], g.spec, pts, ids, g.warp_alpha)
  ref = Z[f'jac/{name}/jacobian']
  assert ref.shape == (pts.shape[0], 3, 3)
  assert float(np.abs(J.numpy() - ref).max()) < 2e-6 * max(1.0, float(np.abs(ref).max()))
  warped = O.warp_field_apply(O.tree_to(g.params, torch.float32)['warp_field'], g.spec, pts.float(), ids, g.warp_alpha)
  np.testing.assert_allclose(warped.numpy(), Z[f'jac/{name}/warped_points'], rtol=0, atol=2e-6)


@pytest.mark.parametrize('name', ['se3_small', 'translation_small', 'pivot_small'])
def test_compute_background_loss(name):
  """training.py:118-135 with the reference run's recorded draws."""
  g = Golden(name)
  loss = O.compute_background_loss(g.params, g.spec, torch.from_numpy(Z[f'bg/{name}/points']),
                                   torch.from_numpy(Z[f'bg/{name}/ids'].astype(np.int64)),
                                   torch.from_numpy(Z[f'bg/{name}/noise']), g.warp_alpha)
  np.testing.assert_allclose(loss.numpy(), Z[f'bg/{name}/loss'], rtol=2e-4, atol=1e-9)


def test_level_regularisers_are_differentiable():
  """The regulariser terms used by the GPU gradient tests: finite, and the elastic term reaches
  the warp parameters through the Jacobian (double backward)."""
  g = Golden('se3_small')
  p64 = O.tree_to(g.params, torch.float64)
  leaves = []

  def mark(t):
    for v in t.values():
      if isinstance(v, dict):
        mark(v)
      else:
        v.requires_grad_(True)
        leaves.append(v)
  mark(p64['warp_field'])
  fwd = O.render_forward(p64, g.spec, g.rays, warp_alpha=g.warp_alpha, dtype=torch.float64, return_points=True)
  r = O.level_regularisers(p64, g.spec, fwd['coarse'], g.rays, g.warp_alpha, use_elastic_loss=True,
                           use_warp_reg_loss=True)
  (r['loss/elastic'] + r['loss/warp_reg']).backward()
  le = float(r["loss/elastic"].detach())
  assert math.isfinite(le) and le > 0
  assert sum(float(v.grad.abs().sum()) for v in leaves if v.grad is not None) > 0
