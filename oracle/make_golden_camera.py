"""Generate tests/golden/camera_*.npz by executing the reference's own Camera class.

TEST INFRASTRUCTURE.  Authoring container only (needs /root/reference):

    python oracle/make_golden_camera.py

nerfies/camera.py is imported unmodified (its only non-numpy dependency,
tf.io.gfile for file IO, is stubbed by oracle/jaxshim/tensorflow).  Each fixture
stores the camera parameters and the outputs of Camera.get_pixel_centers /
pixels_to_rays (camera.py:244-269,317-321) and datasets/core.py:50-75's origins
for a whole small frame plus a set of off-grid pixels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = os.environ.get('NERFIES_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'jaxshim'))
sys.path.insert(0, REFERENCE)
sys.path.insert(0, REPO)

from nerfies import camera as ref_camera  # noqa: E402  (the reference)
from oracle import camera_oracle as C  # noqa: E402

CASES = {
    'camera_distorted': dict(seed=1, width=96, height=54, distortion=True, skew=0.0),
    'camera_pinhole': dict(seed=2, width=64, height=48, distortion=False, skew=0.0),
    'camera_skew_distorted': dict(seed=3, width=80, height=60, distortion=True, skew=0.7),
}


def main():
  out_dir = os.path.join(REPO, 'tests', 'golden')
  for name, kw in CASES.items():
    cam = C.synthetic_camera(**kw)
    ref = ref_camera.Camera(**{k: v for k, v in cam.items()})
    pixels = ref.get_pixel_centers()
    dirs = ref.pixels_to_rays(pixels)
    origins = np.tile(ref.position[None, None, :], ref.image_shape + (1,)).astype(np.float32)
    rng = np.random.RandomState(7)
    off = (rng.uniform(size=(257, 2)) * np.array(cam['image_size'], np.float32)).astype(np.float32)
    off_dirs = ref.pixels_to_rays(off)
    np.savez_compressed(
        os.path.join(out_dir, name + '.npz'),
        **{'cam_' + k: v for k, v in cam.items()},
        pixels=pixels.astype(np.float32), directions=dirs.astype(np.float32), origins=origins,
        off_pixels=off, off_directions=off_dirs.astype(np.float32),
        projected=ref.project(ref.position + dirs.reshape(-1, 3)[::7] * np.float32(1.5)).astype(np.float32))
    print(name, dirs.shape, float(np.abs(np.linalg.norm(dirs, axis=-1) - 1).max()))


if __name__ == '__main__':
  main()
