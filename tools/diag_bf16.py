"""Diagnostic: tensor-core path vs bf16-operand oracle vs fp32 oracle (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nerfies_oracle as O
from tests.golden_util import model_from_spec, rel_err, spec_to_dict, tree_to_device
from tests.test_parity_gpu import _render_level, _bf16_case

for dims in sys.argv[1:] or ['quarterhd']:
  spec, n, alpha = _bf16_case(dims)
  p = O.make_trained_like(O.init_params(spec, 21), seed=22)
  rays = O.synthetic_rays(n, spec, seed=23)
  model = model_from_spec(spec_to_dict(spec), precision='bf16', device='cuda:0', batch_size=n)
  pg = tree_to_device(p, 'cuda:0')
  ref32 = O.render_forward(p, spec, rays, warp_alpha=alpha)
  for lv, level in ((0, 'coarse'), (1, 'fine')):
    z = ref32[level]['z_vals']
    got = _render_level(model, pg, lv, rays, z, alpha)
    with O.bf16_operands():
      ref = O.render_level(p, spec, level, rays, z, alpha)
    r32 = O.render_level(p, spec, level, rays, z, alpha)
    smp = torch.cat([ref['sample_rgb'], ref['sample_sigma'][..., None]], -1)
    smp32 = torch.cat([r32['sample_rgb'], r32['sample_sigma'][..., None]], -1)
    print(dims, level, 'samples: got-vs-bf16oracle %.3e  got-vs-fp32 %.3e  bf16oracle-vs-fp32 %.3e' % (
        rel_err(got['samples'], smp), rel_err(got['samples'], smp32), rel_err(smp, smp32)))
    d = (got['samples'] - smp).abs()
    print('   mean abs diff rgb %.3e sigma %.3e ; oracle-bf16-vs-fp32 mean abs rgb %.3e sigma %.3e' % (
        d[..., :3].mean(), d[..., 3].mean(), (smp - smp32)[..., :3].abs().mean(), (smp - smp32)[..., 3].abs().mean()))
    for k in ['rgb', 'depth', 'acc', 'weights'] + (['warped_points'] if spec.use_warp else []):
      print('   %-14s got-vs-bf16oracle %.3e  got-vs-fp32 %.3e  bf16oracle-vs-fp32 %.3e' % (
          k, rel_err(got[k], ref[k]), rel_err(got[k], r32[k]), rel_err(ref[k], r32[k])))
    # per-row worst
    rows = d[..., 3].reshape(-1)
    i = int(rows.argmax())
    print('   worst sigma row', i, 'got', got['samples'].reshape(-1, 4)[i].tolist(), 'ref', smp.reshape(-1, 4)[i].tolist())
