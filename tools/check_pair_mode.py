"""Bit-identity of the CTA-pair kernel modes against the default tcgen05 kernel, plus a quick rate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nerfies_b200 as nb

mode = sys.argv[1] if len(sys.argv) > 1 else '2'
B = 8192
model, params = nb.construct_nerf(0, bench.model_config(), B, range(200), [0], range(200), 0.02, 0.83,
                                  precision='bf16', device='cuda:0')
rays = bench.synthetic_rays(B, 1)
rays = {'origins': rays['origins'].cuda(), 'directions': rays['directions'].cuda(),
        'metadata': {k: v.cuda() for k, v in rays['metadata'].items()}}
os.environ.pop('NFB_TC_PAIR', None)
a = model.apply({'params': params}, rays, warp_extra={'alpha': 8.0})
torch.cuda.synchronize()
os.environ['NFB_TC_PAIR'] = mode
b = model.apply({'params': params}, rays, warp_extra={'alpha': 8.0})
torch.cuda.synchronize()
same = all(torch.equal(a[lv][k], b[lv][k]) for lv in a for k in ('rgb', 'depth', 'acc'))
t0 = time.time()
for _ in range(5):
  model.apply({'params': params}, rays, warp_extra={'alpha': 8.0})
torch.cuda.synchronize()
dt = (time.time() - t0) / 5
print('pair mode %s: bit-identical=%s, %.1f M ray-samples/s (wall, %d rays)' % (mode, same, B * 384 / dt / 1e6, B))
