"""Timeline of the tensor-core field kernels' block 0 (GPU box).

  python tools/build_variant.py trace -DNFB_TRACE
  PREC=fp16x3|bf16 [STEP_LO=8 STEP_HI=10] python tools/trace_tc.py
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('NFB_LIB_PATH', os.path.join(ROOT, 'nerfies_b200', '_variants', 'libnfb_trace.so'))
import torch
import bench
import nerfies_b200 as nb

PREC = os.environ.get('PREC', 'fp16x3')
wl = bench.WORKLOADS['northstar']
B = 8192
model, params = nb.construct_nerf(0, bench.model_config(wl), B, range(200), range(2), range(200), 0.02, 0.83, precision=PREC, device='cuda:0')
rays = bench.synthetic_rays(B, 1, wl)
rays = {'origins': rays['origins'].cuda(), 'directions': rays['directions'].cuda(), 'metadata': {k: v.cuda() for k, v in rays['metadata'].items()}}
out = model.apply({'params': params}, rays, warp_extra={'alpha': 8.0})
torch.cuda.synchronize()
hd = model.handle(B)
cap = 40000
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
per = cap // 4
recs = []
for role in range(4):
  for i in range(min(t[role], per)):
    tag, clk = t[4 + role * per * 2 + 2 * i], t[4 + role * per * 2 + 2 * i + 1]
    recs.append((role, tag >> 8, tag & 0xff, clk))
recs.sort(key=lambda r: r[3])
names = {0: {0: 'MMA step start', 1: 'MMA chunk0 issued', 2: 'MMA chunk1 issued'},
         1: {0: 'E0 acc0 ready', 1: 'E0 c0 math done', 2: 'E0 x_free', 3: 'E0 c0 stored', 4: 'E0 acc1 ready', 5: 'E0 step done'},
         2: {0: 'E1 acc0 ready', 1: 'E1 c0 math done', 2: 'E1 x_free', 3: 'E1 c0 stored', 4: 'E1 acc1 ready', 5: 'E1 step done'}}
starts = [i for i, r in enumerate(recs) if r[0] == 0 and r[1] == 0 and r[2] == 0]
print('precision', PREC, 'tiles traced:', len(starts), 'records', len(recs))
if len(starts) > 3:
  a, b = starts[2], starts[3]
  base = recs[a][3]
  print('tile period (cycles):', recs[b][3] - recs[a][3])
  # per-step summary: start of the step's first unit -> start of the next step's first unit; waits
  step_start, waits = {}, {}
  last_unit = None
  for r in recs[a:b]:
    if r[0] == 0 and r[2] >= 64:
      step_start.setdefault(r[1], r[3] - base)
      last_unit = r
      if r[2] > 64:
        waits.setdefault(r[1], []).append(r[2] - 64)
  ss = sorted(step_start.items())
  print('step: start cycle, duration, units that had to wait (bit 1 weights, 2/4/8 x_ready[0/1/2])')
  for i, (st, c) in enumerate(ss):
    nxt = ss[i + 1][1] if i + 1 < len(ss) else recs[b][3] - base
    print('  step %2d  start %7d  dur %6d  waits %s' % (st, c, nxt - c, waits.get(st, [])))
  lo, hi = int(os.environ.get('STEP_LO', '9')), int(os.environ.get('STEP_HI', '10'))
  for r in recs[a:b]:
    if lo <= r[1] <= hi:
      if r[0] == 0 and r[2] >= 64: nm = 'MMA  unit begins%s' % ('' if r[2] == 64 else '  WAITS for bits %d' % (r[2] - 64))
      elif r[0] == 3: nm = 'PROD stage free -> copy issued, piece %d' % r[2]
      else: nm = names[r[0]].get(r[2], 'ev %d' % r[2])
      print('%8d  step %2d  %s' % (r[3] - base, r[1], nm))
