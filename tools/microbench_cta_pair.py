"""cta_group::2 MMA rate (nfb_selftest_gemm2 with reps): cycles per M=256 x N x K=16 MMA as seen by
the issuing CTA, next to the single-CTA M=128 figures of tools/microbench.py."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfies_b200 import _lib

lib = _lib.load()
res = {}
for N in (64, 128, 256):
  K = 256
  A = torch.randn(256, K).cuda(); W = (torch.randn(K, N) * 0.1).cuda(); C = torch.empty(256, N, device='cuda')
  out = (ctypes.c_longlong * 2)()
  best = None
  for _ in range(5):
    _lib.check(lib.nfb_selftest_gemm2(K, N, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(W.data_ptr()),
                                      ctypes.c_void_p(C.data_ptr()), 64, out, None))
    cyc = out[0] / out[1]
    best = cyc if best is None else min(best, cyc)
  res['N=%d' % N] = {'cycles_per_mma_M256_K16': round(best, 1), 'mmas': int(out[1]),
                     'flop_per_clk_per_sm': round(2 * 128 * N * 16 / best, 1)}
print(json.dumps({'kernel': 'tc_selftest2_kernel (tcgen05.mma.cta_group::2, M=256)', 'results': res,
                  'note': '1-CTA reference (profiles/r01_microbench.txt): M=128 N=128 64 clk, N=256 128 clk per MMA'}))
