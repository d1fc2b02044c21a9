"""tcgen05.mma with the A operand in tensor memory (nfb_selftest_gemm3): cycles per MMA
(M=128, K=16, kind::f16) for N = 64 / 128 / 256, three-chain fp16x3 schedule."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfies_b200 import _lib

lib = _lib.load()
res = {}
for N in (64, 128, 256):
  K = 192 if N == 256 else 256
  A = torch.randn(128, K).cuda()
  W = (torch.randn(K, N) * 0.1).cuda()
  C = torch.empty(128, N, device='cuda')
  out = (ctypes.c_longlong * 2)()
  for reps in (1, 64):
    _lib.check(lib.nfb_selftest_gemm3(K, N, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(W.data_ptr()),
                                      ctypes.c_void_p(C.data_ptr()), reps, out, None))
  err = float((C.double() - A.double() @ W.double()).abs().max())
  res[f'N{N}'] = {'cycles': out[0], 'mmas': out[1], 'cycles_per_mma': out[0] / out[1], 'max_abs_err': err}
print(json.dumps(res))
