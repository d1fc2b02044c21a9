"""Camera geometry on the evaluation path: rays of a frame generated on the GPU.

Mirror of the part of `nerfies.camera.Camera` (nerfies/camera.py:108-341) and of
`datasets/core.py:50-75 camera_to_rays` that `eval.py` uses per rendered frame:
same constructor arguments, JSON layout, property names and error behaviour;
`pixels_to_rays` / `camera_to_rays` run `camera_rays_kernel` through the C ABI
(`nfb_camera_rays`, `nfb_pixels_to_rays`) and return torch tensors on the device
instead of numpy arrays.  There is no CPU path.
"""
import copy
import ctypes
import json

import numpy as np
import torch

from nerfies_b200 import _lib


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Camera:
  """Class to handle camera geometry (field names of camera.py:110-137)."""

  def __init__(self, orientation, position, focal_length, principal_point, image_size, skew=0.0,
               pixel_aspect_ratio=1.0, radial_distortion=None, tangential_distortion=None,
               dtype=np.float32):
    if dtype != np.float32:
      raise ValueError('nerfies_b200.Camera computes in float32 like the reference default')
    if radial_distortion is None:
      radial_distortion = np.array([0.0, 0.0, 0.0], dtype)
    if tangential_distortion is None:
      tangential_distortion = np.array([0.0, 0.0], dtype)
    self.orientation = np.array(orientation, dtype)
    self.position = np.array(position, dtype)
    self.focal_length = np.array(focal_length, dtype)
    self.principal_point = np.array(principal_point, dtype)
    self.skew = np.array(skew, dtype)
    self.pixel_aspect_ratio = np.array(pixel_aspect_ratio, dtype)
    self.radial_distortion = np.array(radial_distortion, dtype)
    self.tangential_distortion = np.array(tangential_distortion, dtype)
    self.image_size = np.array(image_size, np.uint32)
    self.dtype = dtype
    if self.orientation.shape != (3, 3) or self.position.shape != (3,):
      raise ValueError('orientation must be (3,3) and position (3,)')

  # ---- JSON (camera.py:139-180) ----
  @classmethod
  def from_json(cls, path):
    """Loads a JSON camera into memory."""
    with open(str(path), 'r') as fp:
      camera_json = json.load(fp)
    if 'tangential' in camera_json:          # old camera JSON (camera.py:147-148)
      camera_json['tangential_distortion'] = camera_json['tangential']
    return cls(
        orientation=np.asarray(camera_json['orientation']),
        position=np.asarray(camera_json['position']),
        focal_length=camera_json['focal_length'],
        principal_point=np.asarray(camera_json['principal_point']),
        skew=camera_json['skew'],
        pixel_aspect_ratio=camera_json['pixel_aspect_ratio'],
        radial_distortion=np.asarray(camera_json['radial_distortion']),
        tangential_distortion=np.asarray(camera_json['tangential_distortion']),
        image_size=np.asarray(camera_json['image_size']))

  def to_json(self):
    return {k: (v.tolist() if hasattr(v, 'tolist') else v) for k, v in self.get_parameters().items()}

  def get_parameters(self):
    return {
        'orientation': self.orientation, 'position': self.position,
        'focal_length': self.focal_length, 'principal_point': self.principal_point,
        'skew': self.skew, 'pixel_aspect_ratio': self.pixel_aspect_ratio,
        'radial_distortion': self.radial_distortion,
        'tangential_distortion': self.tangential_distortion, 'image_size': self.image_size,
    }

  # ---- properties (camera.py:182-223) ----
  @property
  def scale_factor_x(self):
    return self.focal_length

  @property
  def scale_factor_y(self):
    return self.focal_length * self.pixel_aspect_ratio

  @property
  def principal_point_x(self):
    return self.principal_point[0]

  @property
  def principal_point_y(self):
    return self.principal_point[1]

  @property
  def has_tangential_distortion(self):
    return any(self.tangential_distortion != 0.0)

  @property
  def has_radial_distortion(self):
    return any(self.radial_distortion != 0.0)

  @property
  def image_size_y(self):
    return self.image_size[1]

  @property
  def image_size_x(self):
    return self.image_size[0]

  @property
  def image_shape(self):
    return int(self.image_size_y), int(self.image_size_x)

  @property
  def optical_axis(self):
    return self.orientation[2, :]

  @property
  def translation(self):
    return -np.matmul(self.orientation, self.position)

  def copy(self):
    return copy.deepcopy(self)

  def scale(self, scale):
    """Scales the camera (camera.py:323-341)."""
    if scale <= 0:
      raise ValueError('scale needs to be positive.')
    return Camera(
        orientation=self.orientation.copy(), position=self.position.copy(),
        focal_length=self.focal_length * scale,
        principal_point=self.principal_point.copy() * scale, skew=self.skew,
        pixel_aspect_ratio=self.pixel_aspect_ratio,
        radial_distortion=self.radial_distortion.copy(),
        tangential_distortion=self.tangential_distortion.copy(),
        image_size=np.array((int(round(self.image_size[0] * scale)),
                             int(round(self.image_size[1] * scale)))))

  # ---- the C-ABI view ----
  def _struct(self):
    c = _lib.NfbCamera()
    c.orientation[:] = [float(v) for v in self.orientation.reshape(-1)]
    c.position[:] = [float(v) for v in self.position]
    c.focal_length = float(self.focal_length)
    c.principal_point[:] = [float(v) for v in self.principal_point]
    c.skew = float(self.skew)
    c.pixel_aspect_ratio = float(self.pixel_aspect_ratio)
    c.radial_distortion[:] = [float(v) for v in self.radial_distortion]
    c.tangential_distortion[:] = [float(v) for v in self.tangential_distortion]
    c.image_size[:] = [int(self.image_size[0]), int(self.image_size[1])]
    return c

  # ---- GPU paths ----
  def pixels_to_rays(self, pixels):
    """Unit world-space ray directions of float32 pixel positions (..., 2) on a CUDA device.

    Replaces Camera.pixels_to_rays (camera.py:244-269); same argument checks."""
    if not torch.is_tensor(pixels):
      raise TypeError('pixels must be a torch tensor on a CUDA device')
    if pixels.shape[-1] != 2:
      raise ValueError('The last dimension of pixels must be 2.')
    if pixels.dtype != torch.float32:
      raise ValueError(f'pixels dtype ({pixels.dtype!r}) must match camera dtype (float32)')
    if not pixels.is_cuda:
      raise ValueError('pixels must live on a CUDA device: nerfies_b200 has no CPU path')
    lib = _lib.load()
    batch_shape = pixels.shape[:-1]
    flat = pixels.reshape(-1, 2).contiguous()
    out = torch.empty(flat.shape[0], 3, dtype=torch.float32, device=pixels.device)
    with torch.cuda.device(pixels.device):
      _lib.check(lib.nfb_pixels_to_rays(ctypes.byref(self._struct()), _ptr(flat), flat.shape[0],
                                        _ptr(out), _stream(pixels.device)))
    return out.reshape(*batch_shape, 3)

  def get_pixel_centers(self, device):
    """Pixel centres (H, W, 2) on `device` (camera.py:317-321)."""
    return camera_to_rays(self, device)['pixels']


def camera_to_rays(camera, device, first_pixel=0, count=None):
  """Rays of a frame, generated on the GPU (datasets/core.py:50-75).

  Returns {'origins', 'directions', 'pixels'}: (H, W, 3|3|2) float32 tensors on
  `device` - or (count, ·) for a row-major pixel sub-range (one rank's slice of a
  frame, one render_image chunk)."""
  device = torch.device(device)
  if device.type != 'cuda':
    raise ValueError('camera_to_rays needs a CUDA device: nerfies_b200 has no CPU path')
  lib = _lib.load()
  h, w = camera.image_shape
  whole = count is None and first_pixel == 0
  n = h * w - first_pixel if count is None else int(count)
  origins = torch.empty(n, 3, dtype=torch.float32, device=device)
  directions = torch.empty(n, 3, dtype=torch.float32, device=device)
  pixels = torch.empty(n, 2, dtype=torch.float32, device=device)
  with torch.cuda.device(device):
    _lib.check(lib.nfb_camera_rays(ctypes.byref(camera._struct()), int(first_pixel), n, _ptr(origins),
                                   _ptr(directions), _ptr(pixels), _stream(device)))
  if whole:
    return {'origins': origins.reshape(h, w, 3), 'directions': directions.reshape(h, w, 3),
            'pixels': pixels.reshape(h, w, 2)}
  return {'origins': origins, 'directions': directions, 'pixels': pixels}
