"""Loading of tests/golden/*.npz (made by oracle/make_golden.py from the
reference's own source) into oracle-side structures."""
import json
import os

import numpy as np
import torch

from oracle import nerfies_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['se3_small', 'translation_small', 'nowarp_variants',
         'alpha_cond_init', 'se3_stratified', 'quarterhd_dims',
         'test_local_dims']


def unflatten(flat):
  tree = {}
  for k, v in flat.items():
    node = tree
    parts = k.split('/')
    for p in parts[:-1]:
      node = node.setdefault(p, {})
    node[parts[-1]] = v
  return tree


def flatten(tree, prefix=''):
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(flatten(v, prefix + k + '/'))
    else:
      out[prefix + k] = v
  return out


class Golden:
  def __init__(self, name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    self.name = name
    self.spec_dict = json.loads(str(z['spec_json']))
    d = dict(self.spec_dict)
    d['nerf_skips'] = tuple(d['nerf_skips'])
    d['warp_skips'] = tuple(d['warp_skips'])
    self.spec = O.OracleSpec(**d)
    self.warp_alpha = float(z['warp_alpha'])
    self.rays = {
        'origins': torch.from_numpy(z['rays/origins']),
        'directions': torch.from_numpy(z['rays/directions']),
        'metadata': {k.split('/')[-1]: torch.from_numpy(
            z[k].astype(np.int32)) for k in z.files
                     if k.startswith('rays/metadata/')},
    }
    if 'oracle_param_seed' in z.files:
      p = O.init_params(self.spec, int(z['oracle_param_seed']))
      tl = int(z['trained_like_seed'])
      if tl >= 0:
        p = O.make_trained_like(p, seed=tl)
      checksum = sum(float(v.double().abs().sum())
                     for v in flatten(p).values())
      ref = float(z['param_checksum'])
      assert abs(checksum - ref) <= 1e-6 * ref, (
          'torch RNG drift: regenerated parameters differ from the fixture')
      self.params = p
    else:
      self.params = unflatten({
          k[len('params/'):]: torch.from_numpy(z[k]) for k in z.files
          if k.startswith('params/')})
    self.out = unflatten({k[len('out/'):]: torch.from_numpy(z[k])
                          for k in z.files if k.startswith('out/')})
    self.t_rand = torch.from_numpy(z['t_rand']) if 't_rand' in z.files else None
    self.u_rand = torch.from_numpy(z['u_rand']) if 'u_rand' in z.files else None
    self.warp = None
    if 'warp/points' in z.files:
      self.warp = {
          'points': torch.from_numpy(z['warp/points']),
          'ids': torch.from_numpy(z['warp/ids'].astype(np.int32)),
          'warped_points': torch.from_numpy(z['warp/warped_points']),
      }


def rel_err(a, b, floor=1e-2):
  """max |a-b| / (|b| + floor)."""
  a = a.double()
  b = b.double()
  return float(((a - b).abs() / (b.abs() + floor)).max())
