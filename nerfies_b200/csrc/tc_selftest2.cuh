// Self-test and micro-benchmark of the CTA-pair (cta_group::2) tcgen05 path that
// round 2 of the field kernel is planned around:
//   C[256][N] = bf16(A[256][K]) x bf16(W[K][N]),  fp32 accumulation,
// executed by a cluster of two CTAs.  CTA r holds rows r*128.. of A and the
// N/2 rows r*N/2.. of the K-major B operand in its own shared memory, at the same
// offsets in both CTAs; the leader CTA's elected thread issues
// tcgen05.mma.cta_group::2 (M = 256), each SM computes its 128 rows of D into its
// own TMEM, and one multicast tcgen05.commit signals both CTAs' barriers.
// Per MMA an SM reads 4 KB of A and N/2 x 32 B of B (2 KB at N = 128) instead of
// 4 KB + 4 KB in the single-CTA form.  Exposed as nfb_selftest_gemm2.
#pragma once
#include "tc_common.cuh"
#include "tc_selftest.cuh"

namespace nfb {
namespace tc {

constexpr int kSelf2SmemBytes = 1024 /*align*/ + 2 * kSelfMaxKb * kABlockBytes + 256;

__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the barrier at this shared-memory offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// grid = 2 (one cluster), 160 threads per CTA.  `reps` repeats the whole K loop
// (accumulating: C = reps x A W) for timing; out[0] = cycles from the first issue
// to completion as seen by the leader, out[1] = number of MMAs issued.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
tc_selftest2_kernel(const float* __restrict__ A, int K, const __nv_bfloat16* __restrict__ Wp,
                    int nkb, int N, float* __restrict__ C, int reps, long long* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_blocks = base;                                   // [kSelfMaxKb][16 KB]: this CTA's 128 rows of A
  uint8_t* b_blocks = a_blocks + kSelfMaxKb * kABlockBytes;   // [kSelfMaxKb][<=16 KB]: this CTA's N/2 rows of B
  uint64_t* acc_ready = reinterpret_cast<uint64_t*>(b_blocks + kSelfMaxKb * kABlockBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int half = N / 2;
  if (tid == 128) {
    mbar_init(acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc2(tmem_slot, 256);
  if (tid < 128) {
    const int r = tid;
    const size_t row = (size_t)rank * 128 + r;
    for (int kb = 0; kb < nkb; ++kb)
      for (int c = 0; c < 8; ++c) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kb * kBlockK + c * 8 + j;
          v[j] = k < K ? A[row * K + k] : 0.f;
        }
        store_chunk(a_blocks + kb * kABlockBytes, r, c, v);
      }
    // B: rows [rank*half, (rank+1)*half) of every packed unit, byte for byte (the
    // chunk swizzle depends on row & 7 only and half is a multiple of 8).
    const int vec_per_kb = half * kRowBytes / 16;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(Wp) +
                                                        ((size_t)kb * N + (size_t)rank * half) * kRowBytes);
      uint4* dst = reinterpret_cast<uint4*>(b_blocks + kb * kABlockBytes);
      for (int i = tid; i < vec_per_kb; i += 128) dst[i] = __ldg(src + i);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  cluster_sync_all();          // both CTAs' operands and barriers are in place
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (rank == 0 && warp == 4) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(256, N);
      const long long t0 = clock64();
      for (int rep = 0; rep < reps; ++rep)
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t a_addr = smem_u32(a_blocks + kb * kABlockBytes);
          const uint32_t b_addr = smem_u32(b_blocks + kb * kABlockBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma2_bf16(tmem_base, make_smem_desc(a_addr + k * 32), make_smem_desc(b_addr + k * 32), idesc,
                       (rep | kb | k) ? 1u : 0u);
        }
      umma2_commit_multicast(acc_ready, 0x3);
      mbar_wait(acc_ready, 0);
      const long long t1 = clock64();
      if (out) {
        out[0] = t1 - t0;
        out[1] = (long long)reps * nkb * 4;
      }
    }
    __syncwarp();
  }
  if (tid < 128) {
    mbar_wait(acc_ready, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const size_t row = (size_t)rank * 128 + tid;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + lane_base + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) C[row * N + c0 + j] = v[j];
    }
    tc_fence_before();
  }
  cluster_sync_all();          // nobody frees TMEM while the peer still reads or computes
  if (warp == 4) tmem_dealloc2(tmem_base, 256);
}

}  // namespace tc
}  // namespace nfb
