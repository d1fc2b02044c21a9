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

__global__ void __launch_bounds__(320, 1)
tc_microbench_kernel(int mode, int n, int reps, int nwarps, long long* out, const uint8_t* gsrc, int smem_words) {
  extern __shared__ __align__(1024) uint8_t raw[];   // 7 x 16 KB: A (2 half-blocks), B stages 1..4, scratch
  uint8_t* a_blk = raw;                       // 16 KB
  uint8_t* b_blk = raw + kABlockBytes;        // up to 32 KB
  __shared__ uint64_t bar, bar2, bar3, bar4;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) stop_flag = 0;
  for (int i = tid; i < smem_words; i += blockDim.x)
  {
    // small bf16 values: constant, or (mode bit 8) pseudo-random mantissas/signs
    uint32_t v = 0x3c003c00u;
    if (mode & 256) {
      uint32_t hsh = (uint32_t)i * 2654435761u;
      hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
      v = (hsh & 0x80ff80ffu) | 0x3c003c00u;
    }
    reinterpret_cast<uint32_t*>(raw)[i] = v;
  }
  mode &= 255;
  if (tid == 256) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); mbar_init(&bar3, 1); mbar_init(&bar4, 1); fence_barrier_init(); }
  if (warp == 8) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (mode == 2) {
    // The fused kernel's issue loop verbatim: whole warp in the loop, per unit a
    // wait on an (already complete) mbarrier + tcgen05.fence::after_thread_sync,
    // then 8 MMAs + commit by one elected lane.  nwarps = variant mask:
    //  bit0: skip the fence, bit1: skip the wait, bit2: commit to the scratch barrier
    //  bit3: B cycles through 4 x 16 KB stages, bit4: warp 9 streams bulk copies
    //  into a scratch stage meanwhile, bit5: warps 0-7 run tcgen05.ld loops meanwhile
    if (warp == 9 && (nwarps & 16)) {
      uint8_t* scratch = raw + 6 * 16384;
      uint32_t ph = 0;
      while (!stop_flag) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar4, 16384);
          bulk_g2s(scratch, gsrc, 16384, &bar4);
        }
        __syncwarp();
        mbar_wait(&bar4, ph);
        ph ^= 1;
      }
    }
    if (warp < 8 && (nwarps & 32)) {
      const uint32_t t_lane = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (warp >> 2) * 256;
      float acc = 0.f;
      while (!stop_flag) {
        float va[32], vb[32];
        tmem_ld32(t_lane + 128, va);
        tmem_ld32(t_lane + 160, vb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += va[j] + vb[j];
      }
      if (acc == 123.456f) out[3] = 1;
    }
    if (warp == 8) {
      const uint32_t idesc = make_idesc_bf16(128, n);
      const uint32_t a = smem_u32(a_blk), b = smem_u32(b_blk);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (!(nwarps & 2)) mbar_wait(&bar3, 1);
        if (!(nwarps & 1)) tc_fence_after();
        const uint32_t b_r = (nwarps & 8) ? b + (r & 3) * 16384 : b;
        if (elect_one()) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + s2 * 256, make_smem_desc(a + s2 * 8192 + k * 32), make_smem_desc(b_r + k * 32), idesc, 1u);
          if (nwarps & 4) umma_commit(&bar2);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (tid == 256) { out[0] = t2 - t0; out[1] = 8LL * reps; out[2] = t1 - t0; }
      stop_flag = 1;
    }
  } else if (mode == 6) {
    // Full replica of the fused kernel's weight ring: warp 9 streams 16 KB units
    // (global -> shared, cp.async.bulk) through kRing stages with full/empty
    // mbarriers; warp 8's elected lane consumes them with issue_unit() exactly as
    // field_tc_kernel does (8 MMAs + commit to empty[stage] per unit).
    constexpr int kRing = 4;
    __shared__ uint64_t rfull[kRing], rempty[kRing];
    if (tid == 0) {
      for (int i = 0; i < kRing; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], 1); }
      fence_barrier_init();
    }
    __syncthreads();
    uint8_t* ring = raw + 160 * 1024;
    const uint32_t unit_bytes = (uint32_t)n * kRowBytes;      // n = MMA N (64 or 128)
    if (warp == 9) {
      if (elect_one()) {
        for (int it = 0; it < reps; ++it) {
          const int sgp = it % kRing;
          mbar_wait(&rempty[sgp], ((it / kRing) & 1) ^ 1);
          mbar_arrive_expect_tx(&rfull[sgp], unit_bytes);
          bulk_g2s(ring + sgp * 16384, gsrc + (size_t)(it & 7) * 16384, unit_bytes, &rfull[sgp]);
        }
      }
      __syncwarp();
    } else if (warp == 8) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, n);
        const uint64_t hi = make_smem_desc(0);
        const uint32_t base_lo = (smem_u32(raw) & 0x3FFFFu) >> 4;
        const uint32_t st_lo = base_lo + (160 * 1024 >> 4);
        const uint32_t bf = smem_u32(&rfull[0]), be = smem_u32(&rempty[0]);
        uint32_t sg = 0, wph = 0, ready = 0;
        const long long t0 = clock64();
        for (int it = 0; it < reps; ++it) {
          if (!(ready & 1)) mbar_wait(&rfull[sg], wph);
          const uint32_t ablk = (it & 3) * 1024;
          const uint64_t bd = hi | (uint64_t)(st_lo + sg * 1024);
          const uint64_t ad0 = hi | (uint64_t)(base_lo + ablk), ad1 = hi | (uint64_t)(base_lo + ablk + 4096);
          const uint32_t nsg = (sg + 1 == kRing) ? 0 : sg + 1;
          const uint32_t nwph = (sg + 1 == kRing) ? wph ^ 1 : wph;
          ready = issue_unit<false>(tmem_base, tmem_base + 256, ad0, ad1, bd, idesc, 1u, be + sg * 8, 0u, 0u,
                             bf + nsg * 8, nwph, 0u, 0u, 0u);
          sg = nsg; wph = nwph;
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t2 - t0; out[1] = 8LL * reps; out[2] = 0;
      }
      __syncwarp();
    }
  } else if (mode >= 3 && mode <= 5) {
    const int layout = mode - 3;
    // The current issuer: one elected-lane region per 4 units, barrier probed
    // between the 6th and 7th MMA of a unit.  Variant bits as in mode 2 (b4 bulk
    // copies, b5 tcgen05.ld from 8 warps) plus b2: smem st/ld traffic from warps 0-7.
    if (warp == 9 && (nwarps & 16)) {
      uint8_t* scratch = raw + 6 * 16384;
      uint32_t ph = 0;
      while (!stop_flag) {
        if (elect_one()) { mbar_arrive_expect_tx(&bar4, 16384); bulk_g2s(scratch, gsrc, 16384, &bar4); }
        __syncwarp();
        mbar_wait(&bar4, ph);
        ph ^= 1;
      }
    }
    if (warp < 8 && (nwarps & (32 | 4))) {
      const uint32_t t_lane = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (warp >> 2) * 256;
      uint8_t* scratch = raw + 5 * 16384;
      float acc = 0.f;
      while (!stop_flag) {
        if (nwarps & 32) {
          float va[32], vb[32];
          tmem_ld32(t_lane + 128, va);
          tmem_ld32(t_lane + 160, vb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += va[j] + vb[j];
        }
        if (nwarps & 4) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            *reinterpret_cast<uint4*>(scratch + swz_off(tid & 127, q)) = make_uint4(tid, q, 0, 0);
            acc += *reinterpret_cast<volatile float*>(scratch + q * 16);
          }
        }
        if (nwarps & 8) {     // the epilogue's hand-over: proxy fence + mbarrier arrive
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&bar2);
        }
      }
      if (acc == 123.456f) out[3] = 1;
    }
    if (warp == 8) {
      const uint32_t idesc = make_idesc_bf16(128, n);
      const uint64_t hi = make_smem_desc(0);
      // layout: 0 = compact (A at 0 / +8 KB, B stages from 16 KB);
      //         1 = the fused kernel's (A blocks at b*16 KB and (4+b)*16 KB, B stages from 160 KB)
      const uint32_t base_lo = (smem_u32(raw) & 0x3FFFFu) >> 4;
      const uint32_t a_lo = base_lo, b_lo = layout ? base_lo + (160 * 1024 >> 4) : base_lo + 1024;
      const uint32_t a1_off = layout ? (64 * 1024 >> 4) : 512;
      const uint32_t dchunk = layout >= 2 ? 128 : 0;
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 4) {
        if (elect_one()) {
          bool ready = mbar_test(&bar3, 1);
          for (int u = 0; u < 4; ++u) {
            if (!ready) mbar_wait(&bar3, 1);
            tc_fence_after();
            const uint64_t bd = hi | (uint64_t)(b_lo + (u & 3) * 1024);
            const uint32_t ablk = layout ? (u & 3) * 1024 : 0;
            const uint64_t ad0 = hi | (uint64_t)(a_lo + ablk), ad1 = hi | (uint64_t)(a_lo + ablk + a1_off);
            const uint32_t dd = tmem_base + (((r >> 2) & 1) ? dchunk : 0);
            umma_bf16(dd, ad0, bd, idesc, 1u);
            umma_bf16(dd, ad0 + 2, bd + 2, idesc, 1u);
            umma_bf16(dd, ad0 + 4, bd + 4, idesc, 1u);
            umma_bf16(dd, ad0 + 6, bd + 6, idesc, 1u);
            umma_bf16(dd + 256, ad1, bd, idesc, 1u);
            umma_bf16(dd + 256, ad1 + 2, bd + 2, idesc, 1u);
            ready = mbar_test(&bar3, 1);
            umma_bf16(dd + 256, ad1 + 4, bd + 4, idesc, 1u);
            umma_bf16(dd + 256, ad1 + 6, bd + 6, idesc, 1u);
            umma_commit(&bar2);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (tid == 256 && blockIdx.x == 0) { out[0] = t2 - t0; out[1] = 8LL * reps; out[2] = t1 - t0; }
      stop_flag = 1;
    }
  } else if (mode == 0) {
    if (tid == 256) {
      const uint32_t idesc = make_idesc_bf16(128, n);
      const uint32_t a = smem_u32(a_blk), b = smem_u32(b_blk);
      // nwarps doubles as a variant mask in mode 0:
      //  bit0: commit (to a scratch barrier) after every 8 MMAs
      //  bit1: alternate the A block / accumulator every 4 MMAs (two sub-tiles)
      //  bit2: 8 other warps generate shared-memory load/store traffic meanwhile
      const int variant = nwarps;
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t a_r = (variant & 2) ? a + (r & 1) * 8192 : a;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + (r & 1) * 256, make_smem_desc(a_r + k * 32), make_smem_desc(b + k * 32), idesc, 1u);
        if ((variant & 1) && (r & 1)) umma_commit(&bar2);
      }
      umma_commit(&bar);
      const long long t1 = clock64();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[0] = t2 - t0; out[1] = 4LL * reps; out[2] = t1 - t0;
      stop_flag = 1;
    } else if (warp < 8 && (nwarps & 4)) {
      // epilogue-like traffic: 128-bit swizzled stores + broadcast loads
      uint8_t* scratch = b_blk + 128 * kRowBytes;   // upper half of the B region (unused for n<=128)
      float acc = 0.f;
      while (!stop_flag) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          *reinterpret_cast<uint4*>(scratch + swz_off(tid & 127, q)) = make_uint4(tid, q, 0, 0);
          acc += *reinterpret_cast<volatile float*>(scratch + q * 16);
        }
      }
      if (acc == 123.456f) out[3] = 1;
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
