// Self-test and micro-benchmark of the A-operand-in-TMEM form of tcgen05.mma that
// the fp16x3 field kernel (field_tc3.cuh) is built on:
//   C[128][N] = A[128][K] x W[K][N] evaluated as three fp16 chains
//       A_hi W_hi + A_lo W_hi + A_hi W_lo      (fp32 accumulation in TMEM)
// with BOTH images of A resident in tensor memory: thread r (TMEM lane r) packs
// its row into fp16 pairs and writes them with tcgen05.st.32x32b; one 32-bit TMEM
// column holds K elements 2c and 2c+1 of the row, so a K = 16 MMA reads 8 columns
// and stepping K adds 8 to the column address.  B comes from shared memory (the
// pre-swizzled [W_hi | W_lo] units of pack_weight_x3_kernel).
// TMEM map (512 columns): accumulator 0..255 | A_hi 256..383 | A_lo 384..511.
// Exposed as nfb_selftest_gemm3; `reps` repeats the chains for timing.
#pragma once
#include "tc_common.cuh"

namespace nfb {
namespace tc {

constexpr int kSelf3MaxKb = 4;                         // K <= 256
constexpr int kSelf3WBytes = 192 * 1024;                // all weight units stay resident: nkb x 2 x n_rows x 128 B
constexpr int kSelf3SmemBytes = 1024 /*align*/ + kSelf3WBytes + 256;
constexpr uint32_t kTmemAHi = 256, kTmemALo = 384;

__device__ __forceinline__ void split_pair_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  float fa, fb;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
      : "=f"(fa), "=f"(fb) : "r"(hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - fb), "f"(a - fa));
}

// Wp: per K-block [W_hi (n_rows x 128 B) | W_lo (n_rows x 128 B)] (pack_weight_x3_kernel).
__global__ void __launch_bounds__(160, 1)
tc_selftest3_kernel(const float* __restrict__ A, int K, const uint8_t* __restrict__ Wp, int nkb, int n_rows,
                    int N, float inv_scale, float* __restrict__ C, int reps, long long* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_all = base;                                        // [nkb][2][n_rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kSelf3WBytes);
  uint64_t* w_full = bars;
  uint64_t* acc_ready = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t unit_bytes = 2u * (uint32_t)n_rows * kRowBytes;
  if (tid == 128) {
    mbar_init(w_full, 1); mbar_init(acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 128) {
    mbar_arrive_expect_tx(w_full, unit_bytes * nkb);
    for (int kb = 0; kb < nkb; ++kb)
      bulk_g2s(w_all + (size_t)kb * unit_bytes, Wp + (size_t)kb * unit_bytes, unit_bytes, w_full);
  }
  if (tid < 128) {
    // A -> TMEM: 32 K columns (16 packed pairs) per store
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < nkb * kBlockK; k0 += 32) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = k0 + 2 * j;
        const float a = k < K ? A[(size_t)tid * K + k] : 0.f;
        const float b = k + 1 < K ? A[(size_t)tid * K + k + 1] : 0.f;
        split_pair_f16(a, b, hi[j], lo[j]);
      }
      tmem_st16(lane_base + kTmemAHi + (k0 >> 1), hi);
      tmem_st16(lane_base + kTmemALo + (k0 >> 1), lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) {
    if (elect_one()) {
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16(128, n_rows);
      const long long t0 = clock64();
      for (int rep = 0; rep < reps; ++rep)
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t b_hi = smem_u32(w_all + (size_t)kb * unit_bytes);
          const uint32_t b_lo = b_hi + (uint32_t)n_rows * kRowBytes;
          const uint32_t a_hi = tmem_base + kTmemAHi + kb * 32, a_lo = tmem_base + kTmemALo + kb * 32;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(tmem_base, a_hi + k * 8, make_smem_desc(b_hi + k * 32), idesc, (rep | kb | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tmem_base, a_lo + k * 8, make_smem_desc(b_hi + k * 32), idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tmem_base, a_hi + k * 8, make_smem_desc(b_lo + k * 32), idesc, 1u);
        }
      umma_commit(acc_ready);
      mbar_wait(acc_ready, 0);
      const long long t1 = clock64();
      if (out) {
        out[0] = t1 - t0;
        out[1] = (long long)reps * nkb * 12;
      }
    }
    __syncwarp();
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
        if (c0 + j < N) C[(size_t)tid * N + c0 + j] = v[j] * inv_scale;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace nfb
