import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfies_b200 import _lib
lib = _lib.load()
torch.zeros(1).cuda()
out = (ctypes.c_longlong * 3)()
for n in (64, 128, 256):
  for reps in (64, 512):
    _lib.check(lib.nfb_selftest_microbench(0, n, reps, 0, out))
    print(f'MMA M=128 N={n} K=16 SS: {out[0]/out[1]:.1f} cycles/MMA over {out[1]} MMAs (issue {out[2]/out[1]:.1f}/MMA)')
for variant in (1, 2, 3, 4, 7):
  _lib.check(lib.nfb_selftest_microbench(0, 128, 512, variant, out))
  print(f'MMA N=128 variant {variant:03b} (bit0 commit/8, bit1 alt A+D, bit2 smem traffic): {out[0]/out[1]:.1f} cycles/MMA')
for variant in (4, 12, 20, 36, 60):
  _lib.check(lib.nfb_selftest_microbench(2, 128, 512, variant, out))
  print(f'issue-loop N=128 variant {variant:06b} (b2 commit/unit, b3 B cycles 4 stages, b4 concurrent bulk copies, b5 concurrent LDTM x8 warps): {out[0]/out[1]:.1f} cycles/MMA')
for variant in (0, 52):
  _lib.check(lib.nfb_selftest_microbench(3, 128, 512, variant, out))
  print(f'probe-ahead issuer N=128 variant {variant:06b} (b2 smem st/ld x8 warps, b3 fence.proxy.async+arrive, b4 bulk copies, b5 LDTM x8 warps): {out[0]/out[1]:.1f} cycles/MMA')
for mode in (3, 3 + 256):
  _lib.check(lib.nfb_selftest_microbench(mode, 128, 512, 0, out))
  print(f'probe-ahead issuer N=128, {"random" if mode & 256 else "constant"} operand data: {out[0]/out[1]:.1f} cycles/MMA')
for mode, name in ((3, 'compact layout'), (4, 'fused-kernel layout (A blocks 16 KB apart, sub-tiles 64 KB apart, B at 160 KB+)'), (5, 'fused layout + alternating accumulator chunks')):
  _lib.check(lib.nfb_selftest_microbench(mode, 128, 512, 0, out))
  print(f'probe-ahead issuer N=128, {name}: {out[0]/out[1]:.1f} cycles/MMA')
for n in (128, 64):
  for grid in (1, 148):
    _lib.check(lib.nfb_selftest_microbench(6 + ((grid - 1) << 9), n, 1024, 0, out))
    print(f'ring replica (4 x 16 KB stages, real cp.async.bulk + full/empty barriers), N={n}, {grid} CTAs: {out[0]/out[1]:.1f} cycles/MMA = {8*out[0]/out[1]:.0f} cycles/unit')
for grid in (1,):
  _lib.check(lib.nfb_selftest_microbench(3 + ((grid - 1) << 9), 128, 2048, 0, out))
  print(f'probe-ahead issuer N=128 on {grid} CTAs/SMs concurrently: {out[0]/out[1]:.1f} cycles/MMA (block 0)')
for mode in (0, 256):
  _lib.check(lib.nfb_selftest_microbench(mode, 256, 512, 0, out))
  print(f'MMA chain N=256, {"random" if mode & 256 else "constant"} operand data: {out[0]/out[1]:.1f} cycles/MMA')
for nw in (1, 4, 8):
  _lib.check(lib.nfb_selftest_microbench(1, 128, 256, nw, out))
  per = out[0] / out[1]
  print(f'LDTM 32x32b.x32 by {nw} warps: {per:.1f} cycles per 4 KB load per warp -> {nw*4096/per:.0f} B/clk/SM')
