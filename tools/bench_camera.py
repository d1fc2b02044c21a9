"""Times nfb_camera_rays (camera -> rays on the GPU) against its HBM roofline and the
numpy oracle on the host.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import nerfies_b200 as nb
from oracle import camera_oracle as C

res = {}
# distorted cameras are bound by the 10 Newton steps of the undistortion; pinhole cameras
# (no distortion: camera.py:201-207 skips the solve) are the HBM-bound case.
for name, (w, h, dist) in {'1080p': (1920, 1080, True), '8k': (7680, 4320, True),
                            '1080p_pinhole': (1920, 1080, False), '8k_pinhole': (7680, 4320, False)}.items():
  cam = C.synthetic_camera(3, w, h, distortion=dist, skew=0.1 if dist else 0.0)
  c = nb.camera.Camera(**cam)
  for _ in range(3):
    nb.camera.camera_to_rays(c, 'cuda:0')
  torch.cuda.synchronize()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
  reps = 20
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  total = 0.0
  for _ in range(reps):
    flush.zero_()
    ev[0].record()
    nb.camera.camera_to_rays(c, 'cuda:0')
    ev[1].record()
    torch.cuda.synchronize()
    total += ev[0].elapsed_time(ev[1])
  ms = total / reps          # includes the three torch.empty allocations (cached allocator)
  bytes_written = w * h * 32
  res[name] = {'pixels': w * h, 'ms': ms, 'GB_per_s': bytes_written / ms / 1e6,
               'Mpixels_per_s': w * h / ms / 1e3}
  if name == '1080p':
    t0 = time.time()
    C.camera_to_rays(cam)
    res[name]['numpy_oracle_ms'] = (time.time() - t0) * 1e3
peak = None
try:
  peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
except Exception:
  pass
print(json.dumps({'kernel': 'camera_rays_kernel', 'algorithmic_bytes_per_pixel': 32, 'results': res,
                  'measured_peaks': peak}))
