"""Closed-form known-answer tests for the oracle, each derived
from the cited reference lines."""
import math

import numpy as np
import torch

from oracle import nerfies_oracle as O


def test_sinusoidal_encoder_order_and_cos_identity():
  # modules.py:213-228: identity, then per frequency [sin block, "cos" block].
  x = torch.tensor([1.0, 2.0, 3.0])
  got = O.sinusoidal_encode(x, 2)
  hp = np.float32(np.pi / 2)
  a = np.array([1, 2, 3], np.float32)
  exp = np.concatenate([a, np.sin(a), np.sin(a + hp), np.sin(2 * a),
                        np.sin(2 * a + hp)]).astype(np.float32)
  assert got.shape == (15,)
  np.testing.assert_allclose(got.numpy(), exp, rtol=0, atol=2e-7)
  # zero frequencies -> identity (modules.py:210-211).
  assert torch.equal(O.sinusoidal_encode(x, 0), x)


def test_cos_is_sin_of_shifted_float32_angle():
  # At 2^9 x the shifted-sine differs measurably from a true cosine: the kernel
  # must reproduce sin(fl32(a + fl32(pi/2))).
  x = torch.tensor([[1.37, -0.81, 0.22]])
  enc = O.sinusoidal_encode(x, 10)
  a = (x.numpy().astype(np.float32) * np.float32(512.0))
  shifted = np.sin((a + np.float32(np.pi / 2)).astype(np.float32))
  np.testing.assert_allclose(enc[0, 3 + 9 * 6 + 3:3 + 9 * 6 + 6].numpy(),
                             shifted[0], atol=2e-7)
  true_cos = np.cos(a.astype(np.float64))[0]
  assert np.abs(shifted[0] - true_cos).max() > 2e-6


def test_cosine_easing_window():
  # modules.py:274-294.
  np.testing.assert_allclose(O.cosine_easing_window(6, 0.0), np.zeros(6),
                             atol=1e-7)
  np.testing.assert_allclose(O.cosine_easing_window(6, 6.0), np.ones(6),
                             atol=1e-7)
  np.testing.assert_allclose(O.cosine_easing_window(5, 2.5),
                             [1, 1, 0.5, 0, 0], atol=1e-6)


def test_exp_se3_pure_rotation_about_z():
  # rigid_body.py:54-89.
  S = torch.tensor([0., 0., 1., 0., 0., 0.])
  R, p = O.exp_se3(S, torch.tensor(math.pi / 2))
  np.testing.assert_allclose(R.numpy(), [[0, -1, 0], [1, 0, 0], [0, 0, 1]],
                             atol=1e-6)
  np.testing.assert_allclose(p.numpy(), [0, 0, 0], atol=1e-7)


def test_exp_se3_general_matches_closed_form():
  g = torch.Generator().manual_seed(0)
  w = torch.randn(5, 3, generator=g, dtype=torch.float64)
  w = w / w.norm(dim=-1, keepdim=True)
  v = torch.randn(5, 3, generator=g, dtype=torch.float64)
  th = torch.rand(5, generator=g, dtype=torch.float64) * 3
  R, p = O.exp_se3(torch.cat([w, v], -1), th)
  for i in range(5):
    W = np.array([[0, -w[i, 2], w[i, 1]], [w[i, 2], 0, -w[i, 0]],
                  [-w[i, 1], w[i, 0], 0]])
    t = float(th[i])
    Rr = np.eye(3) + math.sin(t) * W + (1 - math.cos(t)) * W @ W
    pr = (t * np.eye(3) + (1 - math.cos(t)) * W + (t - math.sin(t)) * W @ W
          ) @ v[i].numpy()
    np.testing.assert_allclose(R[i].numpy(), Rr, atol=1e-12)
    np.testing.assert_allclose(p[i].numpy(), pr, atol=1e-12)
    # R is a rotation.
    np.testing.assert_allclose(R[i].numpy() @ R[i].numpy().T, np.eye(3),
                               atol=1e-12)


def _render(sigma, z=None, white=False, infinity=True):
  B, S = sigma.shape
  if z is None:
    z = torch.linspace(1.0, 2.0, S).expand(B, S)
  rgb = torch.full((B, S, 3), 0.25)
  dirs = torch.tensor([[0., 0., 1.]]).expand(B, 3)
  return O.volumetric_rendering(rgb, sigma, z, dirs, white, infinity), z


def test_volumetric_rendering_zero_density():
  # model_utils.py:104-133: weights 0 -> rgb 0 (1 on white), acc/depth 0.
  out, _ = _render(torch.zeros(2, 8))
  for k in ('rgb', 'depth', 'med_depth', 'acc', 'weights'):
    assert float(out[k].abs().max()) == 0.0
  out, _ = _render(torch.zeros(2, 8), white=True)
  np.testing.assert_allclose(out['rgb'].numpy(), 1.0)


def test_volumetric_rendering_single_opaque_sample():
  sigma = torch.zeros(1, 8)
  sigma[0, 3] = 1e6
  out, z = _render(sigma)
  w = out['weights'][0]
  assert abs(float(w[3]) - 1.0) < 1e-6 and float(w.sum() - w[3]) < 1e-6
  assert abs(float(out['depth'][0]) - float(z[0, 3])) < 1e-6
  assert float(out['med_depth'][0]) == float(z[0, 3])
  assert abs(float(out['acc'][0]) - 1.0) < 1e-6
  np.testing.assert_allclose(out['rgb'][0].numpy(), 0.25, atol=1e-6)


def test_sample_at_infinity_moves_last_weight_out_of_acc():
  # model_utils.py:121-126: acc excludes the last (infinite) sample, white
  # background uses the pre-override acc.
  sigma = torch.full((1, 4), 0.5)
  out, _ = _render(sigma, infinity=True)
  w = out['weights'][0]
  assert abs(float(w.sum()) - 1.0) < 1e-6           # last alpha is 1
  assert abs(float(out['acc'][0]) - float(w[:-1].sum())) < 1e-7
  out2, _ = _render(sigma, infinity=False)
  assert float(out2['weights'][0, -1]) < 1e-12       # 1e-19 last interval


def test_piecewise_constant_pdf_uniform_weights():
  # model_utils.py:153-187: uniform pdf + u=linspace -> linspace over bins.
  bins = torch.linspace(2.0, 6.0, 9)[None]
  w = torch.ones(1, 8)
  z = O.piecewise_constant_pdf(bins, w, 17)
  np.testing.assert_allclose(z[0].numpy(), np.linspace(2.0, 6.0, 17),
                             atol=2e-6)


def test_inverse_cdf_equals_right_searchsorted():
  # model_utils.py:169-179 == idx=searchsorted(cdf,u,'right')-1 with clamps.
  g = torch.Generator().manual_seed(3)
  for zero_heavy in (False, True):
    w = torch.rand(4, 14, generator=g)
    if zero_heavy:
      w = w * (torch.rand(4, 14, generator=g) > 0.7)
    bins = torch.sort(torch.rand(4, 15, generator=g), dim=-1).values
    for u_rand in (None, torch.rand(4, 20, generator=g)):
      z = O.piecewise_constant_pdf(bins, w, 20, u_rand)
      ww = w + 1e-5
      pdf = ww / ww.sum(-1, keepdim=True)
      cdf = torch.cat([torch.zeros(4, 1), torch.cumsum(pdf, -1)], -1)
      u = u_rand if u_rand is not None else torch.from_numpy(
          np.linspace(0., 1., 20, dtype=np.float32)).expand(4, 20)
      idx = torch.searchsorted(cdf, u.contiguous(), right=True) - 1
      lo = idx.clamp(0, 13)
      hi = (idx + 1).clamp(1, 14)
      c0, c1 = cdf.gather(-1, lo), cdf.gather(-1, hi)
      b0, b1 = bins.gather(-1, lo), bins.gather(-1, hi)
      den = c1 - c0
      den = torch.where(den < 1e-5, torch.ones_like(den), den)
      ref = b0 + (u - c0) / den * (b1 - b0)
      assert torch.equal(z, ref)


def test_sample_along_rays_deterministic():
  # model_utils.py:56-70.
  o = torch.zeros(3, 3)
  d = torch.tensor([[0., 0., 1.]]).expand(3, 3)
  z, pts = O.sample_along_rays(o, d, 5, 2.0, 6.0, False)
  np.testing.assert_allclose(z.numpy(), np.tile([2, 3, 4, 5, 6], (3, 1)),
                             atol=1e-6)
  np.testing.assert_allclose(pts[..., 2].numpy(), z.numpy())
  z, _ = O.sample_along_rays(o, d, 3, 1.0, 4.0, True)
  np.testing.assert_allclose(z[0].numpy(), [1.0, 1.6, 4.0], atol=1e-6)


def test_glo_encoder_squeezes_trailing_dim():
  # glo.py:50-53.
  p = {'embed': {'embedding': torch.arange(12.).reshape(4, 3)}}
  ids = torch.tensor([[2], [0]], dtype=torch.int32)
  assert torch.equal(O.glo_encode(p, ids), p['embed']['embedding'][[2, 0]])
  assert torch.equal(O.glo_encode(p, ids[:, 0]), p['embed']['embedding'][[2, 0]])


def test_rgb_condition_quirk_and_param_counts():
  # models.py:206-207: appearance code reaches rgb only with
  # use_alpha_condition; SURVEY §8a R7: 589,700 params/level @Fp=8, cond 27.
  spec = O.OracleSpec(num_nerf_point_freqs=8, use_appearance_metadata=True,
                      use_rgb_condition=True, use_warp=True,
                      num_warp_embeddings=1)
  assert O.cond_dims(spec) == (0, 0, 27)
  p = O.init_params(spec, 0)
  count = lambda t: sum(count(v) for v in t.values()) if isinstance(
      t, dict) else t.numel()
  assert count(p['nerf_mlps_coarse']) == 589700
  wf = dict(p['warp_field'])
  wf.pop('metadata_encoder')
  assert count(wf) == 98566
  heads = p['warp_field']['branches_w']['logit']['kernel']
  assert float(heads.min()) >= 0 and float(heads.max()) < 1e-4
