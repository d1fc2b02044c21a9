"""CPU restatement of the reference's camera -> rays path.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
may import this.  numpy float32, operation for operation as the reference writes
it (nerfies/camera.py is itself numpy), so the restatement is expected to agree
with the reference bit for bit; tests/golden/camera_*.npz hold outputs of the
reference's own Camera class (oracle/make_golden_camera.py) and pin it.

  undistort            nerfies/camera.py:26-105
  pixel_to_local_rays  nerfies/camera.py:225-242
  pixels_to_rays       nerfies/camera.py:244-269
  get_pixel_centers    nerfies/camera.py:317-321
  camera_to_rays       nerfies/datasets/core.py:50-75
  project              nerfies/camera.py:283-315 (used by the round-trip property test)
"""
import numpy as np

F = np.float32


def make_camera(orientation, position, focal_length, principal_point, image_size, skew=0.0,
                pixel_aspect_ratio=1.0, radial_distortion=None, tangential_distortion=None):
  """Field names and dtypes of Camera.__init__ (camera.py:110-137)."""
  return {
      'orientation': np.array(orientation, F).reshape(3, 3),
      'position': np.array(position, F).reshape(3),
      'focal_length': np.array(focal_length, F),
      'principal_point': np.array(principal_point, F).reshape(2),
      'skew': np.array(skew, F),
      'pixel_aspect_ratio': np.array(pixel_aspect_ratio, F),
      'radial_distortion': np.array([0, 0, 0] if radial_distortion is None else radial_distortion, F),
      'tangential_distortion': np.array([0, 0] if tangential_distortion is None else tangential_distortion, F),
      'image_size': np.array(image_size, np.uint32).reshape(2),
  }


def _residual_and_jacobian(x, y, xd, yd, k1, k2, k3, p1, p2):
  """camera.py:26-71."""
  r = x * x + y * y
  d = 1.0 + r * (k1 + r * (k2 + k3 * r))
  fx = d * x + 2 * p1 * x * y + p2 * (r + 2 * x * x) - xd
  fy = d * y + 2 * p2 * x * y + p1 * (r + 2 * y * y) - yd
  d_r = (k1 + r * (2.0 * k2 + 3.0 * k3 * r))
  d_x = 2.0 * x * d_r
  d_y = 2.0 * y * d_r
  fx_x = d + d_x * x + 2.0 * p1 * y + 6.0 * p2 * x
  fx_y = d_y * x + 2.0 * p1 * x + 2.0 * p2 * y
  fy_x = d_x * y + 2.0 * p2 * y + 2.0 * p1 * x
  fy_y = d + d_y * y + 2.0 * p2 * x + 6.0 * p1 * y
  return fx, fy, fx_x, fx_y, fy_x, fy_y


def undistort(xd, yd, k1, k2, k3, p1, p2, eps=1e-9, max_iterations=10):
  """camera.py:74-105: ten Newton steps from the distorted point, no early exit."""
  x = xd.copy()
  y = yd.copy()
  for _ in range(max_iterations):
    fx, fy, fx_x, fx_y, fy_x, fy_y = _residual_and_jacobian(x, y, xd, yd, k1, k2, k3, p1, p2)
    denominator = fy_x * fx_y - fx_x * fy_y
    x_numerator = fx * fy_y - fy * fx_y
    y_numerator = fy * fx_x - fx * fy_x
    ok = np.abs(denominator) > eps
    with np.errstate(divide='ignore', invalid='ignore'):
      step_x = np.where(ok, x_numerator / denominator, np.zeros_like(denominator))
      step_y = np.where(ok, y_numerator / denominator, np.zeros_like(denominator))
    x = x + step_x
    y = y + step_y
  return x, y


def pixel_to_local_rays(cam, pixels):
  """camera.py:225-242."""
  scale_factor_x = cam['focal_length']
  scale_factor_y = cam['focal_length'] * cam['pixel_aspect_ratio']
  y = (pixels[..., 1] - cam['principal_point'][1]) / scale_factor_y
  x = (pixels[..., 0] - cam['principal_point'][0] - y * cam['skew']) / scale_factor_x
  rd, td = cam['radial_distortion'], cam['tangential_distortion']
  if any(rd != 0.0) or any(td != 0.0):
    x, y = undistort(x, y, k1=rd[0], k2=rd[1], k3=rd[2], p1=td[0], p2=td[1])
  dirs = np.stack([x, y, np.ones_like(x)], axis=-1)
  return dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)


def pixels_to_rays(cam, pixels):
  """camera.py:244-269 (world-space unit directions)."""
  if pixels.shape[-1] != 2:
    raise ValueError('The last dimension of pixels must be 2.')
  if pixels.dtype != F:
    raise ValueError('pixels dtype must be float32')
  batch_shape = pixels.shape[:-1]
  pixels = np.reshape(pixels, (-1, 2))
  local = pixel_to_local_rays(cam, pixels)
  rays_dir = np.matmul(cam['orientation'].T, local[..., np.newaxis])
  rays_dir = np.squeeze(rays_dir, axis=-1)
  rays_dir /= np.linalg.norm(rays_dir, axis=-1, keepdims=True)
  return rays_dir.reshape((*batch_shape, 3))


def get_pixel_centers(cam):
  """camera.py:317-321."""
  w, h = int(cam['image_size'][0]), int(cam['image_size'][1])
  xx, yy = np.meshgrid(np.arange(w, dtype=F), np.arange(h, dtype=F))
  return np.stack([xx, yy], axis=-1) + 0.5


def camera_to_rays(cam):
  """datasets/core.py:50-75."""
  h, w = int(cam['image_size'][1]), int(cam['image_size'][0])
  origins = np.tile(cam['position'][None, None, :], (h, w, 1))
  pixels = get_pixel_centers(cam)
  return {'origins': origins.astype(F), 'directions': pixels_to_rays(cam, pixels).astype(F),
          'pixels': pixels.astype(F)}


def project(cam, points):
  """camera.py:283-315."""
  batch_shape = points.shape[:-1]
  points = points.reshape((-1, 3))
  local = (np.matmul(cam['orientation'], (points - cam['position']).T)).T
  x = local[..., 0] / local[..., 2]
  y = local[..., 1] / local[..., 2]
  r2 = x**2 + y**2
  rd, td = cam['radial_distortion'], cam['tangential_distortion']
  distortion = 1.0 + r2 * (rd[0] + r2 * (rd[1] + rd[2] * r2))
  x_times_y = x * y
  xn = x * distortion + 2.0 * td[0] * x_times_y + td[1] * (r2 + 2.0 * x**2)
  yn = y * distortion + 2.0 * td[1] * x_times_y + td[0] * (r2 + 2.0 * y**2)
  px = cam['focal_length'] * xn + cam['skew'] * yn + cam['principal_point'][0]
  py = cam['focal_length'] * cam['pixel_aspect_ratio'] * yn + cam['principal_point'][1]
  return np.stack([px, py], axis=-1).reshape((*batch_shape, 2))


def synthetic_camera(seed, width, height, distortion=True, skew=0.0):
  """A plausible capture camera (phone-like intrinsics, random pose)."""
  rng = np.random.RandomState(seed)
  a = rng.normal(size=(3, 3))
  q, r = np.linalg.qr(a)
  q = q * np.sign(np.diag(r))
  if np.linalg.det(q) < 0:
    q[2] = -q[2]
  return make_camera(
      orientation=q, position=rng.normal(size=3) * 0.3, focal_length=0.9 * width,
      principal_point=[width / 2 + rng.normal() * 3, height / 2 + rng.normal() * 3],
      image_size=[width, height], skew=skew, pixel_aspect_ratio=1.0 + 0.01 * rng.normal(),
      radial_distortion=[0.05, -0.08, 0.02] if distortion else None,
      tangential_distortion=[1e-3, -5e-4] if distortion else None)
