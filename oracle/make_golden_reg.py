"""Generate tests/golden/regularisers.npz by executing the reference's own Python source.

TEST INFRASTRUCTURE.  Authoring container only (needs /root/reference, read-only):

    python oracle/make_golden_reg.py

Known answers for the training-tier regularisers, produced by the
UNMODIFIED reference functions running on oracle/jaxshim (numpy float32 primitives):
  * training.compute_elastic_loss (training.py:71-115) for every loss type on a set of
    Jacobians (near identity, sheared, one with a negative determinant);
  * utils.general_loss_with_squared_residual (utils.py:264-331) over alpha / scale;
  * model_utils.compute_depth_index (model_utils.py:242-245);
  * warp_field.apply(..., return_jacobian=True) (warping.py:385-387) for the SE(3) field, the
    SE(3) field with pivot + translation and the translation field, with the parameters of the
    existing fixtures - through the shim's NUMERICAL jax.jacfwd (float64 central differences of
    the reference's own `warp`; see oracle/jaxshim/jax/__init__.py);
  * training.compute_background_loss (training.py:118-135) with its random draws recorded.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (sets up the shim + reference import paths)

import jax  # noqa: E402  (the shim)
from jax import random as jrandom  # noqa: E402
from nerfies import configs, model_utils, models, training, utils  # noqa: E402  (the reference)

CASES = {
    'se3_small': (dict(use_warp=True, warp_field_type='se3', num_nerf_point_freqs=8, num_warp_freqs=8,
                       use_appearance_metadata=True, use_camera_metadata=True, warp_kwargs={'trunk_width': 32}),
                  dict(n_app=5, n_cam=2, n_warp=7, near=0.02, far=0.83, num_rays=12, seed=11)),
    'translation_small': (dict(use_warp=True, warp_field_type='translation', num_nerf_point_freqs=6,
                               num_warp_freqs=6, num_warp_features=4, use_appearance_metadata=True,
                               warp_kwargs={'hidden_channels': 32}),
                          dict(n_app=3, n_cam=1, n_warp=4, near=0.05, far=1.2, num_rays=10, seed=12)),
    'pivot_small': (dict(use_warp=True, warp_field_type='se3', num_nerf_point_freqs=8, num_warp_freqs=8,
                         use_appearance_metadata=True,
                         warp_kwargs={'trunk_width': 32, 'use_pivot': True, 'use_translation': True}),
                    dict(n_app=4, n_cam=1, n_warp=6, near=0.02, far=0.83, num_rays=8, seed=21)),
}
SMALL = dict(nerf_trunk_width=64, nerf_rgb_branch_width=32, num_coarse_samples=16, num_fine_samples=16)

# jax.random.choice is not in the shim: uniform choice from the numpy generator, recorded.
_CHOICE, _NORMAL = [], []


def _choice(key, a, shape=()):
  idx = key.integers(0, len(a), size=tuple(shape))
  out = np.asarray(a)[idx]
  _CHOICE.append(out)
  return out


_orig_normal = jrandom.normal


def _normal(key, shape, *a, **k):
  n = _orig_normal(key, shape, *a, **k)
  _NORMAL.append(n)
  return n


jrandom.choice = _choice
jrandom.normal = _normal


class _State:
  def __init__(self, warp_extra):
    self.warp_extra = warp_extra


def main():
  blob = {}
  rng = np.random.default_rng(2024)
  # ---- elastic loss ----
  J = np.eye(3, dtype=np.float32)[None] + 0.08 * rng.standard_normal((24, 3, 3)).astype(np.float32)
  J[20] = np.diag([1.3, 0.7, 1.05]).astype(np.float32) @ J[20]
  J[21] = J[21] * 2.5
  J[22] = np.diag([1.0, 1.0, -1.0]).astype(np.float32) @ J[22]       # negative determinant (log_det clamp)
  J[23] = np.eye(3, dtype=np.float32)
  blob['elastic/J'] = J
  for t in ('log_svals', 'svals', 'jtj', 'div', 'det', 'log_det', 'nr'):
    res = [training.compute_elastic_loss(J[i], loss_type=t) for i in range(len(J))]
    blob[f'elastic/{t}/loss'] = np.array([np.float32(r[0]) for r in res])
    blob[f'elastic/{t}/residual'] = np.array([np.float32(r[1]) for r in res])
  # ---- general loss ----
  sq = np.concatenate([[0.0], np.logspace(-8, 1, 28)]).astype(np.float32)
  blob['gl/sq'] = sq
  combos = [(-2.0, 0.03), (-2.0, 0.001), (0.0, 0.05), (1.0, 0.1), (2.0, 0.3), (-np.inf, 0.2), (np.inf, 2.0), (-0.5, 0.01)]
  blob['gl/alpha'] = np.array([c[0] for c in combos], np.float32)
  blob['gl/scale'] = np.array([c[1] for c in combos], np.float32)
  blob['gl/loss'] = np.stack([np.asarray(utils.general_loss_with_squared_residual(
      jax.numpy.array(sq), np.float32(a), np.float32(s)), np.float32) for a, s in combos])
  # ---- depth index ----
  w = rng.random((14, 24)).astype(np.float32) ** 6
  w[3] *= 0.01                                                    # never reaches 0.5
  w[5, 0] = 0.9                                                   # first sample already opaque
  blob['depth/weights'] = w
  blob['depth/index'] = np.asarray(model_utils.compute_depth_index(jax.numpy.array(w)), np.int32)
  # ---- warp Jacobians + background loss with the parameters of the existing fixtures ----
  for name, (cfg_kwargs, kw) in CASES.items():
    z = np.load(os.path.join(REPO, 'tests', 'golden', name + '.npz'))
    params = MG.unflatten({k[len('params/'):]: z[k] for k in z.files if k.startswith('params/')})
    cfg = configs.ModelConfig(use_stratified_sampling=False, activation=MG.ACT['relu'],
                              sigma_activation=MG.ACT['softplus'], **{**SMALL, **cfg_kwargs})
    model, _ = models.construct_nerf(
        jax.random.PRNGKey(kw['seed']), cfg, batch_size=kw['num_rays'], appearance_ids=list(range(kw['n_app'])),
        camera_ids=list(range(kw['n_cam'])), warp_ids=list(range(kw['n_warp'])), near=kw['near'], far=kw['far'])
    warp_extra = {'alpha': float(z['warp_alpha']), 'time_alpha': 0.0}
    wf = model.create_warp_field(model, num_batch_dims=1)
    r2 = np.random.default_rng(kw['seed'] + 50)
    pts = (r2.random((20, 3)) * 0.8 - 0.4).astype(np.float32)
    ids = r2.integers(0, kw['n_warp'], size=(20, 1)).astype(np.uint32)
    out = wf.apply({'params': params['warp_field']}, pts, ids, warp_extra, True, False)
    blob[f'jac/{name}/points'] = pts
    blob[f'jac/{name}/ids'] = ids
    blob[f'jac/{name}/warped_points'] = np.asarray(out['warped_points'], np.float32)
    blob[f'jac/{name}/jacobian'] = np.asarray(out['jacobian'], np.float64)
    # compute_background_loss (training.py:118-135): draws recorded
    del _CHOICE[:], _NORMAL[:]
    bpts = (r2.random((16, 3)) * 0.6 - 0.3).astype(np.float32)
    loss = training.compute_background_loss(model, _State(warp_extra), params, jax.random.PRNGKey(kw['seed'] + 9),
                                            bpts, np.float32(0.001))
    blob[f'bg/{name}/points'] = bpts
    blob[f'bg/{name}/ids'] = np.asarray(_CHOICE[0], np.uint32)
    blob[f'bg/{name}/noise'] = np.asarray(np.float32(0.001) * _NORMAL[0], np.float32)
    blob[f'bg/{name}/loss'] = np.asarray(loss, np.float32)
    print(name, 'jacobian[0] =', np.asarray(out['jacobian'])[0].round(4).tolist(), 'bg loss', float(np.mean(loss)))
  path = os.path.join(REPO, 'tests', 'golden', 'regularisers.npz')
  np.savez_compressed(path, **blob)
  print(f'regularisers: {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
  main()
