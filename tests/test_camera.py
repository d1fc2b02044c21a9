"""Camera -> rays: oracle vs the reference's own Camera class
(golden fixtures made by oracle/make_golden_camera.py), host-side mirror logic,
and - on the GPU - the CUDA kernel vs the oracle and vs the fixtures.

Tolerance: directions are unit vectors computed in float32 by ~60 operations; the
CUDA kernel follows the reference's operation order, the only freedom being the
3x3 matmul's accumulation order inside numpy.  Asserted: 2e-6 absolute per
component (north-star tolerance for floating point is 1e-4 relative)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import camera_oracle as C

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = ['camera_distorted', 'camera_pinhole', 'camera_skew_distorted']
TOL = 2e-6


def _load(name):
  z = np.load(os.path.join(GOLDEN, name + '.npz'))
  cam = {k[4:]: z[k] for k in z.files if k.startswith('cam_')}
  return cam, z


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_camera_class(name):
  cam, z = _load(name)
  np.testing.assert_array_equal(C.get_pixel_centers(cam), z['pixels'])
  rays = C.camera_to_rays(cam)
  np.testing.assert_array_equal(rays['origins'], z['origins'])
  # same numpy operations in the same order: bit-exact
  np.testing.assert_array_equal(rays['directions'], z['directions'])
  np.testing.assert_array_equal(C.pixels_to_rays(cam, z['off_pixels']), z['off_directions'])
  pts = cam['position'] + z['directions'].reshape(-1, 3)[::7] * np.float32(1.5)
  np.testing.assert_array_equal(C.project(cam, pts), z['projected'])


def test_oracle_known_answers():
  # identity pose, no distortion: the principal point looks down +z; a pixel one
  # focal length to the right is at 45 degrees.
  cam = C.make_camera(np.eye(3), [1, 2, 3], 100.0, [50.0, 40.0], [100, 80])
  d = C.pixels_to_rays(cam, np.array([[50.0, 40.0], [150.0, 40.0], [50.0, 140.0]], np.float32))
  s = np.float32(np.sqrt(0.5))
  np.testing.assert_allclose(d, [[0, 0, 1], [s, 0, s], [0, s, s]], atol=1e-7)
  rays = C.camera_to_rays(cam)
  assert rays['origins'].shape == (80, 100, 3) and (rays['origins'] == np.float32([1, 2, 3])).all()
  assert rays['pixels'][0, 0].tolist() == [0.5, 0.5] and rays['pixels'][79, 99].tolist() == [99.5, 79.5]
  # distortion round trip: undistort inverts the forward model of project()
  cam = C.synthetic_camera(5, 320, 240)
  px = C.get_pixel_centers(cam)[::9, ::7]
  d = C.pixels_to_rays(cam, px)
  back = C.project(cam, cam['position'] + d.reshape(-1, 3) * np.float32(2.0)).reshape(px.shape)
  assert np.abs(back - px).max() < 2e-3        # pixels


def test_host_camera_mirror(tmp_path):
  import nerfies_b200 as nb
  cam = C.synthetic_camera(4, 64, 48)
  c = nb.camera.Camera(**cam)
  path = tmp_path / 'cam.json'
  path.write_text(json.dumps(c.to_json()))
  c2 = nb.camera.Camera.from_json(path)
  for k, v in c.get_parameters().items():
    np.testing.assert_array_equal(v, c2.get_parameters()[k])
  assert c.image_shape == (48, 64) and c.has_radial_distortion and c.has_tangential_distortion
  half = c.scale(0.5)
  assert half.image_shape == (24, 32) and float(half.focal_length) == float(c.focal_length) * 0.5
  with pytest.raises(ValueError):
    c.scale(0)
  # old JSON key (camera.py:147-148)
  j = c.to_json(); j['tangential'] = j.pop('tangential_distortion')
  path.write_text(json.dumps(j))
  np.testing.assert_array_equal(nb.camera.Camera.from_json(path).tangential_distortion, c.tangential_distortion)
  # argument checks of pixels_to_rays (camera.py:253-257) and the no-CPU-path rule
  with pytest.raises(ValueError):
    c.pixels_to_rays(torch.zeros(4, 3))
  with pytest.raises(ValueError):
    c.pixels_to_rays(torch.zeros(4, 2, dtype=torch.float64))
  with pytest.raises(ValueError):
    c.pixels_to_rays(torch.zeros(4, 2))
  with pytest.raises(ValueError):
    nb.camera.camera_to_rays(c, 'cpu')
  s = c._struct()
  assert list(s.image_size) == [64, 48] and abs(s.focal_length - float(c.focal_length)) == 0


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_rays_match_reference_fixtures(name):
  import nerfies_b200 as nb
  cam, z = _load(name)
  c = nb.camera.Camera(**cam)
  rays = nb.camera.camera_to_rays(c, 'cuda:0')
  assert rays['directions'].shape == z['directions'].shape
  np.testing.assert_array_equal(rays['pixels'].cpu().numpy(), z['pixels'])
  np.testing.assert_array_equal(rays['origins'].cpu().numpy(), z['origins'])
  err = np.abs(rays['directions'].cpu().numpy() - z['directions']).max()
  assert err <= TOL, err
  off = c.pixels_to_rays(torch.from_numpy(z['off_pixels']).cuda())
  assert np.abs(off.cpu().numpy() - z['off_directions']).max() <= TOL


@pytest.mark.gpu
def test_cuda_rays_full_hd_properties_and_ranges():
  """1920x1080 with distortion: vs the oracle on a strided subset, unit norm,
  sub-range generation == slices of the whole frame, project() round trip."""
  import nerfies_b200 as nb
  cam = C.synthetic_camera(11, 1920, 1080, distortion=True, skew=0.3)
  c = nb.camera.Camera(**cam)
  rays = nb.camera.camera_to_rays(c, 'cuda:0')
  d = rays['directions']
  assert d.shape == (1080, 1920, 3)
  assert float((d.norm(dim=-1) - 1).abs().max()) < 3e-7
  sub_px = C.get_pixel_centers(cam)[::37, ::41]
  want = C.pixels_to_rays(cam, sub_px)
  assert np.abs(d[::37, ::41].cpu().numpy() - want).max() <= TOL
  flat = d.reshape(-1, 3)
  part = nb.camera.camera_to_rays(c, 'cuda:0', first_pixel=1234567, count=8192)
  assert torch.equal(part['directions'], flat[1234567:1234567 + 8192])
  assert torch.equal(part['pixels'], rays['pixels'].reshape(-1, 2)[1234567:1234567 + 8192])
  tail = nb.camera.camera_to_rays(c, 'cuda:0', first_pixel=1920 * 1080 - 5)
  assert tail['directions'].shape == (5, 3) and torch.equal(tail['directions'], flat[-5:])
  empty = nb.camera.camera_to_rays(c, 'cuda:0', first_pixel=0, count=0)
  assert empty['directions'].shape == (0, 3)
  with pytest.raises(nb._lib.NfbError):
    nb.camera.camera_to_rays(c, 'cuda:0', first_pixel=1920 * 1080 - 5, count=6)
  pts = cam['position'] + d[::29, ::31].reshape(-1, 3).cpu().numpy() * np.float32(1.7)
  back = C.project(cam, pts)
  assert np.abs(back - rays['pixels'][::29, ::31].reshape(-1, 2).cpu().numpy()).max() < 5e-3


@pytest.mark.gpu
def test_render_image_from_gpu_generated_rays():
  """eval.py:330-348 with the rays made on the GPU: camera_to_rays -> render_image
  equals rendering the oracle's numpy rays of the same camera."""
  import nerfies_b200 as nb
  from oracle import nerfies_oracle as O
  from tests.golden_util import model_from_spec, spec_to_dict, tree_to_device
  spec = O.OracleSpec(num_coarse_samples=32, num_fine_samples=32, near=0.02, far=0.83,
                      num_nerf_point_freqs=8, sigma_activation='softplus', use_warp=True,
                      use_appearance_metadata=True, num_warp_embeddings=4, num_appearance_embeddings=4)
  params = tree_to_device(O.make_trained_like(O.init_params(spec, 3)), 'cuda:0')
  model = model_from_spec(spec_to_dict(spec), precision='fp32', device='cuda:0', batch_size=512)
  cam = C.synthetic_camera(21, 40, 30)
  cam['position'] = (cam['position'] * 0.1).astype(np.float32)
  c = nb.camera.Camera(**cam)
  rays = nb.camera.camera_to_rays(c, 'cuda:0')
  h, w = c.image_shape
  meta = {'warp': torch.full((h, w, 1), 2, dtype=torch.int32, device='cuda:0'),
          'appearance': torch.full((h, w, 1), 1, dtype=torch.int32, device='cuda:0')}
  batch = {'origins': rays['origins'], 'directions': rays['directions'], 'metadata': meta}
  from nerfies_b200.model_utils import Optimizer, TrainState
  model_fn = nb.evaluation.make_model_fn(model)
  state = TrainState(Optimizer({'model': params}), warp_alpha=4.0)
  out = nb.evaluation.render_image(state, batch, model_fn, 1, None, chunk=512)
  ref = C.camera_to_rays(cam)
  batch2 = {'origins': torch.from_numpy(ref['origins']).cuda(), 'directions': torch.from_numpy(ref['directions']).cuda(),
            'metadata': meta}
  out2 = nb.evaluation.render_image(state, batch2, model_fn, 1, None, chunk=512)
  assert out['rgb'].shape == (h, w, 3)
  assert float((out['rgb'] - out2['rgb']).abs().max()) < 1e-4
