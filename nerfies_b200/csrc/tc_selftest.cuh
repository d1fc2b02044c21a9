// Self-test of the tcgen05 plumbing used by the tensor-core field kernel:
// C[128][N] = bf16(A[128][K]) x bf16(W[K][N]) with fp32 accumulation, built from
// exactly the primitives of tc_common.cuh (swizzled operand stores, pre-packed
// weight units moved by cp.async.bulk through a 2-stage mbarrier ring,
// tcgen05.mma into TMEM, tcgen05.ld epilogue).  Exposed as nfb_selftest_gemm.
#pragma once
#include "tc_common.cuh"

namespace nfb {
namespace tc {

constexpr int kSelfMaxKb = 5;
constexpr int kSelfStageBytes = 256 * kRowBytes;     // 32 KB
constexpr int kSelfSmemBytes = 1024 /*align*/ + kSelfMaxKb * kABlockBytes + 2 * kSelfStageBytes + 256;

__global__ void __launch_bounds__(160, 1)
tc_selftest_kernel(const float* __restrict__ A, int K, const __nv_bfloat16* __restrict__ Wp,
                   int nkb, int n_rows, int N, float* __restrict__ C) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_blocks = base;
  uint8_t* w_stage = a_blocks + kSelfMaxKb * kABlockBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_stage + 2 * kSelfStageBytes);
  uint64_t* full = bars;          // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* acc_ready = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 128) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
    mbar_init(acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  if (tid < 128) {
    const int r = tid;
    for (int kb = 0; kb < nkb; ++kb)
      for (int c = 0; c < 8; ++c) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kb * kBlockK + c * 8 + j;
          v[j] = k < K ? A[(size_t)r * K + k] : 0.f;
        }
        store_chunk(a_blocks + kb * kABlockBytes, r, c, v);
      }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t unit_bytes = (uint32_t)n_rows * kRowBytes;

  if (tid == 128) {
    const uint32_t idesc = make_idesc_bf16(128, n_rows);
    for (int kb = 0; kb < nkb && kb < 2; ++kb) {
      mbar_arrive_expect_tx(&full[kb], unit_bytes);
      bulk_g2s(w_stage + kb * kSelfStageBytes, reinterpret_cast<const uint8_t*>(Wp) + (size_t)kb * unit_bytes,
               unit_bytes, &full[kb]);
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb & 1;
      const uint32_t ph = (kb >> 1) & 1;
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(a_blocks + kb * kABlockBytes);
      const uint32_t b_addr = smem_u32(w_stage + st * kSelfStageBytes);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, make_smem_desc(a_addr + k * 32), make_smem_desc(b_addr + k * 32), idesc,
                  (kb | k) ? 1u : 0u);
      umma_commit(&empty[st]);
      if (kb + 2 < nkb) {
        mbar_wait(&empty[st], ph);
        mbar_arrive_expect_tx(&full[st], unit_bytes);
        bulk_g2s(w_stage + st * kSelfStageBytes,
                 reinterpret_cast<const uint8_t*>(Wp) + (size_t)(kb + 2) * unit_bytes, unit_bytes, &full[st]);
      }
    }
    umma_commit(acc_ready);
  }
  if (tid < 128) {
    mbar_wait(acc_ready, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < n_rows; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + lane_base + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < N) C[(size_t)tid * N + c0 + j] = v[j];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
  (void)lane;
}

}  // namespace tc
}  // namespace nfb

// ---------------------------------------------------------------------------
// Micro-benchmarks of the two rates that bound the fused kernel:
//   mode 0: a chain of `reps` x 4 tcgen05.mma (M=128, N=n, K=16, SS operands
//           resident in shared memory) - cycles from first issue to completion;
//   mode 1: `reps` x (two tcgen05.ld 32x32b.x32 + wait) by 4 or 8 warps -
//           cycles seen by warp 0.
// out[0] = cycles, out[1] = work items (MMAs or 32-column loads per warp).
// ---------------------------------------------------------------------------
namespace nfb {
namespace tc {

__global__ void __launch_bounds__(288, 1)
tc_microbench_kernel(int mode, int n, int reps, int nwarps, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* a_blk = raw;                       // 16 KB
  uint8_t* b_blk = raw + kABlockBytes;        // up to 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kABlockBytes + 256 * kRowBytes) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw)[i] = 0x3c003c00u;   // small bf16 values
  if (tid == 256) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 8) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (mode == 0) {
    if (tid == 256) {
      const uint32_t idesc = make_idesc_bf16(128, n);
      const uint32_t a = smem_u32(a_blk), b = smem_u32(b_blk);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + (r & 1) * 256, make_smem_desc(a + k * 32), make_smem_desc(b + k * 32), idesc, 1u);
      }
      umma_commit(&bar);
      const long long t1 = clock64();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[0] = t2 - t0; out[1] = 4LL * reps; out[2] = t1 - t0;
    }
  } else {
    if (warp < nwarps) {
      const uint32_t t_lane = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (warp >> 2) * 256;
      float acc = 0.f;
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        float va[32], vb[32];
        tmem_ld32(t_lane + (r & 3) * 64, va);
        tmem_ld32(t_lane + (r & 3) * 64 + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += va[j] + vb[j];
      }
      const long long t1 = clock64();
      if (tid == 0) { out[0] = t1 - t0; out[1] = 2LL * reps; }
      if (acc == 123.456f) out[3] = 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace nfb
