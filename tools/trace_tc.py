"""Timeline of the tensor-core field kernel's block 0 (GPU box).

Needs the tracer build:  python tools/build_variant.py trace -DNFB_TRACE
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('NFB_LIB_PATH', os.path.join(ROOT, 'nerfies_b200', '_variants', 'libnfb_trace.so'))
import torch
import bench
import nerfies_b200 as nb

B = 8192
model, params = nb.construct_nerf(0, bench.model_config(), B, range(200), [0], range(200), 0.02, 0.83, precision='bf16', device='cuda:0')
rays = bench.synthetic_rays(B, 1)
rays = {'origins': rays['origins'].cuda(), 'directions': rays['directions'].cuda(), 'metadata': {k: v.cuda() for k, v in rays['metadata'].items()}}
out = model.apply({'params': params}, rays, warp_extra={'alpha': 8.0})
torch.cuda.synchronize()
hd = model.handle(B)
cap = 20000
buf = torch.zeros(4 + 2 * cap, dtype=torch.int64, device='cuda')
_lib_check = hd.lib.nfb_set_trace(hd.h, ctypes.c_void_p(buf.data_ptr()), cap)
assert _lib_check == 0, 'tracer build missing: python tools/build_variant.py trace -DNFB_TRACE'
from nerfies_b200 import _lib
from nerfies_b200.models import _ptr, _stream
z = torch.empty(B, 128, device='cuda')
_lib.check(hd.lib.nfb_coarse_z_vals(hd.h, B, None, _ptr(z), _stream()))
o6 = torch.empty(B, 6, device='cuda'); w = torch.empty(B, 128, device='cuda')
ids = rays['metadata']['warp'][:, 0].contiguous(); ida = rays['metadata']['appearance'][:, 0].contiguous()
_lib.check(hd.lib.nfb_render_samples(hd.h, 0, B, 128, _ptr(z), _ptr(rays['origins']), _ptr(rays['directions']), None, _ptr(ids), _ptr(ida), None, 8.0, 0, _ptr(o6), _ptr(w), None, None, _stream()))
torch.cuda.synchronize()
hd.lib.nfb_set_trace(hd.h, None, 0)
t = buf.cpu().tolist()
o = t[4 + 2 * cap - 10:4 + 2 * cap]
if o[0]:
  n = o[3] + o[5] + o[7]
  print('issuer accounting: wall %.0f cyc/unit over %d units; waits+probe-miss %.0f/unit;' % (o[0] / n, n, o[1] / n))
  for name, i in (('N=128', 0), ('N=64', 1), ('N=16', 2)):
    if o[3 + 2 * i]: print('   %s units: %d, inside issue block %.0f cycles each' % (name, o[3 + 2 * i], o[2 + 2 * i] / o[3 + 2 * i]))
per = cap // 4
recs = []
for role in range(4):
  for i in range(min(t[role], per)):
    tag, clk = t[4 + role * per * 2 + 2 * i], t[4 + role * per * 2 + 2 * i + 1]
    if role == 0 and (tag & 0xff) == 40:
      recs.append((role, tag >> 8, 40, recs[-1][3] + 1, clk))
    else:
      recs.append((role, tag >> 8, tag & 0xff, clk, 0))
n = len(recs)
t0 = min(r[3] for r in recs)
recs.sort(key=lambda r: r[3])
names = {0: {0: 'MMA step start', 1: 'MMA chunk0 issued', 2: 'MMA chunk1 issued'},
         1: {0: 'E0 acc0 ready', 1: 'E0 c0 math done', 2: 'E0 x_free', 3: 'E0 c0 stored', 4: 'E0 acc1 ready', 5: 'E0 step done'},
         2: {0: 'E1 acc0 ready', 1: 'E1 c0 math done', 2: 'E1 x_free', 3: 'E1 c0 stored', 4: 'E1 acc1 ready', 5: 'E1 step done'}}
# print the second tile pair (steady state): records 2nd occurrence of step 0 start onwards, ~130 lines
starts = [i for i, r in enumerate(recs) if r[0] == 0 and r[1] == 0 and r[2] == 0]
print('pairs traced:', len(starts), 'records', n)
if len(starts) > 2:
  a, b = starts[1], starts[2]
  base = recs[a][3]
  print('pair period (cycles):', recs[b][3] - recs[a][3])
  lo, hi = int(os.environ.get('STEP_LO', '8')), int(os.environ.get('STEP_HI', '10'))
  for r in recs[a:b]:
    if lo <= r[1] <= hi:
      if r[0] == 0 and r[2] == 40:
        print('          step %2d  MMA  weight waits in segment: %d cycles, %d misses' % (r[1], r[4] >> 8, r[4] & 0xff)); continue
      if r[0] == 0 and 20 <= r[2] < 30: nm = 'MMA  segment begin c%d seg%d' % ((r[2] - 20) // 2, (r[2] - 20) % 2)
      elif r[0] == 0 and 30 <= r[2] < 40: nm = 'MMA  segment issued c%d seg%d' % ((r[2] - 30) // 2, (r[2] - 30) % 2)
      elif r[0] == 0 and r[2] in (50, 51): nm = 'MMA  unit issue begins (weights %s)' % ('ready' if r[2] == 51 else 'NOT ready -> blocking wait')
      elif r[0] == 0 and r[2] in (52, 53): nm = 'MMA  unit issued (next weights %s)' % ('ready' if r[2] == 53 else 'not yet')
      elif r[0] == 3: nm = 'PROD stage free -> copy issued, unit %d' % r[2]
      elif r[0] == 0 and r[2] >= 30: nm = 'MMA  got weights c%d kb%d' % ((r[2] - 30) // 8, (r[2] - 30) % 8)
      elif r[0] == 0 and r[2] >= 10: nm = 'MMA  wait weights c%d kb%d' % ((r[2] - 10) // 8, (r[2] - 10) % 8)
      else: nm = names[r[0]][r[2]]
      print('%8d  step %2d  %s' % (r[3] - base, r[1], nm))
