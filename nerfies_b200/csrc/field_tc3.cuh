// Fused per-sample field evaluation on the 5th-gen tensor cores at fp32-level
// accuracy (NFB_PREC_FP16X3): every Dense layer is evaluated as THREE chains of
// tcgen05.mma (kind::f16, fp16 operands, fp32 accumulation in TMEM) into the same
// accumulator,
//     x W  ~=  x_hi W_hi + x_lo W_hi + x_hi W_lo,
// where v_hi = fp16(v) and v_lo = fp16(v - v_hi).  fp16 keeps 11 significant bits,
// so hi + lo carries 22 bits (relative residual <= 2^-23, absolute floor 2^-25 from
// fp16 subnormals) and the dropped x_lo W_lo term is <= 2^-22 |x||W|: the operand
// error is of the order of one fp32 rounding per product, below the fp32
// accumulation error of a K = 256 dot product.  (A bf16 split would leave 2^-17 per
// operand - not enough for the 1e-4 gate after 8-11 chained layers - and
// kind::tf32 x 3 costs twice the tensor time and 4-byte operands.)  The reference
// evaluates these layers in fp32 (modules.py:39-62, 107-169).
//
// Layout: ONE 128-row tile per CTA at a time.  The ACTIVATIONS LIVE IN TENSOR MEMORY:
// tcgen05.mma takes its A operand from TMEM ([d], [a], b-desc form), so the epilogue
// writes the fp16 hi / lo images of a layer's output with tcgen05.st (thread r = TMEM
// lane r = row r; one 32-bit column holds K elements 2c, 2c+1) and they never touch
// shared memory:
//   TMEM (512 columns): accumulators 0..255 (chunk 0 | chunk 1) | A_hi 256..383 | A_lo 384..511
//   shared memory     : input block hi | lo (2 x 16 KB, the encoded points / conditions,
//                       SS-form MMAs) | weight ring 6 x 32 KB | scratch | barriers
// (The first version kept both images in shared memory - 128 KB - which left room for a
// 2-slot weight ring only: every 256-wide layer stalled ~4 times on the ~1,500-cycle
// latency of a bulk copy, 10.2 K cycles per layer against 6.1 K of MMA work; it also
// put the A reads and the epilogue's swizzled stores on the shared-memory port the
// weights come through.)
// The 3x tensor work per layer hides the epilogue without a second tile: chunk 0's
// epilogue overlaps chunk 1's MMAs, chunk 1's epilogue overlaps the first K-blocks of
// the next layer (those blocks come from chunk 0).
// Warp roles (384 threads): warps 0-7 epilogue, TWO threads per row (warps w and
// w + 4 share a TMEM lane quarter and split a chunk's columns), warp 8 issues the
// MMAs, warp 9 streams the weights (cp.async.bulk, [W_hi | W_lo] per K-block and
// chunk), warps 10-11 complete the control warpgroup (setmaxnreg 40 / 232).
// One issuer "unit" = one K-block of one chunk = 12 MMAs = one ring slot (K-step-major:
// x_hi W_hi with the A collector filled, x_hi W_lo on the collected operand, x_lo W_hi).
// Units of a 256-wide layer are issued in the balanced phase order of x3_unit_order
// (field_tc.cuh); biases reach the epilogue through L1, not the constant bank.
// Measured (profiles/r02_*): 112.8 K cycles per 128-row tile against a 64.2 K tensor-pipe
// floor, tensor pipe active 64.6 %, and the whole chip at its 1000 W power limit
// (~1.72 GHz of 1.965): 0.86 of the power-limited tensor rate in executed FLOPs.
// Build options for A/B timing: NFB_X3_OPT (epilogue scheduling bits), NFB_X3_CLUSTER,
// NFB_X3_NO_COLLECT, NFB_X3_ORDER_OLD, NFB_X3_WARP_ARRIVE, NFB_X3_EXP_* (tools/ab_x3.sh).
#pragma once
#include <cuda_fp16.h>

#include "field_tc.cuh"

namespace nfb {
namespace tc3 {

using namespace nfb::tc;

constexpr int kX3Threads = 384;
constexpr int kX3EpiThreads = 256;
// The weight ring: a slot holds [W_hi | W_lo] of one unit (one bulk copy, one "full"
// barrier, ONE tcgen05.commit to release it: a commit costs ~120 cycles of tensor-pipe
// issue, so it is paid once per 12 MMAs).
constexpr int kX3Slots = 6;
constexpr int kX3SlotBytes = 2 * kStageBytes;                  // 32 KB
constexpr int kInhOff = 0;                                     // input block, hi image
constexpr int kInlOff = kABlockBytes;                          // input block, lo image
constexpr int kStageOff = 2 * kABlockBytes;
constexpr int kPartOff = kStageOff + kX3Slots * kX3SlotBytes;  // alpha partial of the row's second thread
constexpr int kScanOff = kPartOff + 512;                       // fused composite: cross-warp partials (40 floats)
constexpr int kBarOff = kScanOff + 256;
constexpr int kX3SmemBytes = kBarOff + 256;
static_assert(kX3SmemBytes <= 232448, "shared memory");
static_assert(kX3Slots % 2 == 0, "slots are released in pairs");
// TMEM columns
constexpr uint32_t kTmCols = 512, kTmAHi = 256, kTmALo = 384;
// Unit table (TcUnit::a0_lo / a1_lo, built by build_tc_program): bit 31 set = the A
// K-block is in TMEM and bits 0..15 are its column offset; otherwise the value is the
// shared-memory byte offset >> 4 of the input block image.
constexpr uint32_t kUnitATmem = 0x80000000u;
// x_ready arrivals.  NFB_X3_WARP_ARRIVE (A/B build): one arrival per epilogue WARP (bar.warp.sync, then
// lane 0) instead of one per thread - 8 instead of 256 serialised updates of the same barrier word.
#ifdef NFB_X3_WARP_ARRIVE
constexpr int kXrCount = kX3EpiThreads / 32;
__device__ __forceinline__ void xr_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
#else
constexpr int kXrCount = kX3EpiThreads;
__device__ __forceinline__ void xr_arrive(uint64_t* bar) { mbar_arrive(bar); }
#endif

struct X3Bars {
  uint64_t full[kX3Slots];
  uint64_t empty[kX3Slots / 2];  // one per PAIR of slots: released by one commit after the pair's second unit
  uint64_t acc_ready[2];
  uint64_t x_free;
  uint64_t x_ready[3];           // [0] chunk-0 epilogue done | [1] chunk-1 accumulator read | [2] chunk-1 epilogue done
  uint32_t tmem_slot;
};
static_assert(sizeof(X3Bars) <= 256, "barrier block");

// (a, b) -> packed fp16 pairs hi = rn(a, b) and lo = rn(a - hi_a, b - hi_b); `a` is
// the lower half (the lower K column).  Saturating: |v| > 65504 does not become inf.
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  float fa, fb;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
      : "=f"(fa), "=f"(fb) : "r"(hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - fb), "f"(a - fa));
}

// 8 consecutive K columns (one 16-byte chunk) of `row` into the hi and lo images.
__device__ __forceinline__ void store_chunk_x3(uint8_t* bh, uint8_t* bl, int row, int chunk, const float* v) {
  uint4 qh, ql;
  split_pair(v[0], v[1], qh.x, ql.x);
  split_pair(v[2], v[3], qh.y, ql.y);
  split_pair(v[4], v[5], qh.z, ql.z);
  split_pair(v[6], v[7], qh.w, ql.w);
  const uint32_t off = swz_off(row, chunk);
  *reinterpret_cast<uint4*>(bh + off) = qh;
  *reinterpret_cast<uint4*>(bl + off) = ql;
}

// ---------------------------------------------------------------------------
// Positional encoding for the parity mode (modules.py:213-228, 262-271).  The
// reference evaluates sin(2^f x) and sin(fl32(2^f x + fl32(pi/2))) in fp32; 2^f x is
// exact.  Instead of 6 F libm sinf calls per row (the dominant serial section of a
// tile: ~11 K cycles) the octaves are generated by the double-angle recurrence
//     sin 2a = 2 sin a cos a,   cos 2a = 1 - 2 sin^2 a
// in fp64 from an fp64 sin/cos of the base angle: the error doubles per octave from
// ~1e-16, i.e. < 1e-13 at 2^9 - far below half an fp32 ulp, so the features are the
// correctly rounded sines (libm sinf / numpy are within 1-2 ulp of the same values).
// The reference's "cosine" is NOT cos(a): its argument is rounded to fp32 first
// (|error| up to half an ulp of a, 3e-5 at a = 768), which is reproduced exactly:
//     t = fl32(a + hp),  e = t - (a + hp)  (exact in fp64),
//     sin t = sin(a + hp) cos e + cos(a + hp) sin e.
// B200's fp64 pipe runs at half the fp32 rate; this costs ~400 fp64 operations per
// thread instead of ~1000 fp32 instructions with local-memory traffic.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sincos_f64(double x, double& s, double& c) {
  // Cody-Waite reduction by pi/2 (|x| < 2^20 keeps k * PIO2_HI exact enough), then
  // Taylor polynomials on [-pi/4, pi/4] (truncation < 5e-17).
  const double kInvPio2 = 0.63661977236758134308;
  const double kPio2Hi = 1.57079632679489655800e+00, kPio2Lo = 6.12323399573676603587e-17;
  const double kd = rint(x * kInvPio2);
  const int k = (int)kd;
  double r = fma(-kd, kPio2Hi, x);
  r = fma(-kd, kPio2Lo, r);
  const double r2 = r * r;
  double ps = 2.81145725434552076320e-15;             // 1/17!
  ps = fma(ps, r2, -7.64716373181981647590e-13);      // -1/15!
  ps = fma(ps, r2, 1.60590438368216145994e-10);       // 1/13!
  ps = fma(ps, r2, -2.50521083854417187751e-08);      // -1/11!
  ps = fma(ps, r2, 2.75573192239858906526e-06);       // 1/9!
  ps = fma(ps, r2, -1.98412698412698412698e-04);      // -1/7!
  ps = fma(ps, r2, 8.33333333333333333333e-03);       // 1/5!
  ps = fma(ps, r2, -1.66666666666666666667e-01);      // -1/3!
  const double sr = fma(ps * r2, r, r);
  double pc = -1.56192069685862264622e-16;            // -1/18!
  pc = fma(pc, r2, 4.77947733238738529744e-14);       // 1/16!
  pc = fma(pc, r2, -1.14707455977297247139e-11);      // -1/14!
  pc = fma(pc, r2, 2.08767569878680989792e-09);       // 1/12!
  pc = fma(pc, r2, -2.75573192239858906526e-07);      // -1/10!
  pc = fma(pc, r2, 2.48015873015873015873e-05);       // 1/8!
  pc = fma(pc, r2, -1.38888888888888888889e-03);      // -1/6!
  pc = fma(pc, r2, 4.16666666666666666667e-02);       // 1/4!
  pc = fma(pc, r2, -0.5);
  const double cr = fma(pc, r2, 1.0);
  switch (k & 3) {
    case 0: s = sr; c = cr; break;
    case 1: s = cr; c = -sr; break;
    case 2: s = -sr; c = -cr; break;
    default: s = -cr; c = sr; break;
  }
}

// kHS: which 32-column half of the 64-column input block this thread writes.
// posenc_pack_x3 leaves the 32 columns as packed fp16 hi / lo pairs in registers (4 x 16-byte chunks
// per image): the next tile's encoding is computed while the input block is still being read and
// stored later (store_packed_x3).
template <int kHS>
__device__ __forceinline__ void posenc_pack_x3(uint4* qh, uint4* ql, const float* x, int F,
                                               const float* __restrict__ window,
                                               const float* __restrict__ extra, int n_extra) {
  float feat[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) feat[j] = 0.f;
  double sn[3], cs[3];
  float a32[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (kHS == 0) feat[c] = x[c];
    sincos_f64((double)x[c], sn[c], cs[c]);
    a32[c] = x[c];
  }
#pragma unroll
  for (int f = 0; f < 10; ++f) {
    if (f < F) {
      const float w = window ? __ldg(window + f) : 1.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int ks = 3 + f * 6 + c, kc = ks + 3;           // columns of the sin / "cos" feature
        if (ks >= 32 * kHS && ks < 32 * kHS + 32) {
          const float v = (float)sn[c];
          feat[ks - 32 * kHS] = window ? w * v : v;
        }
        if (kc >= 32 * kHS && kc < 32 * kHS + 32) {
          // the reference's fp32 argument t = fl32(a + fl32(pi/2)) = a + pi/2 + eps (|eps| <= ulp(t)/2 +
          // 4.4e-8 <= 4e-5): sin t = cos(a + eps) = cs (1 - eps^2/2) - sn eps  (the eps^3 term is < 1e-14)
          const float t = a32[c] + kHalfPiF;
          const double eps = ((double)t - (double)a32[c]) - 1.57079632679489661923;
          const double v64 = fma(-sn[c], eps, fma(cs[c], -0.5 * eps * eps, cs[c]));
          const float v = (float)v64;
          feat[kc - 32 * kHS] = window ? w * v : v;
        }
        const double s2 = 2.0 * sn[c] * cs[c];
        const double c2 = fma(-2.0 * sn[c], sn[c], 1.0);
        sn[c] = s2; cs[c] = c2;
        a32[c] = a32[c] * 2.0f;                              // exact
      }
    }
  }
  const int d = 3 + 6 * F;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int k = 32 * kHS + j;
    if (k >= d && k < d + n_extra) feat[j] = __ldg(extra + (k - d));
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float* v = feat + 8 * c;
    split_pair(v[0], v[1], qh[c].x, ql[c].x);
    split_pair(v[2], v[3], qh[c].y, ql[c].y);
    split_pair(v[4], v[5], qh[c].z, ql[c].z);
    split_pair(v[6], v[7], qh[c].w, ql[c].w);
  }
}
__device__ __forceinline__ void store_packed_x3(uint8_t* bh, uint8_t* bl, int r, int hs, const uint4* qh, const uint4* ql) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t off = swz_off(r, 4 * hs + c);
    *reinterpret_cast<uint4*>(bh + off) = qh[c];
    *reinterpret_cast<uint4*>(bl + off) = ql[c];
  }
}
template <int kHS>
__device__ __forceinline__ void posenc_block_x3(uint8_t* bh, uint8_t* bl, int r, const float* x, int F,
                                                const float* __restrict__ window,
                                                const float* __restrict__ extra, int n_extra) {
  uint4 qh[4], ql[4];
  posenc_pack_x3<kHS>(qh, ql, x, F, window, extra, n_extra);
  store_packed_x3(bh, bl, r, kHS, qh, ql);
}

__device__ __forceinline__ void cond_to_block_x3(uint8_t* bh, uint8_t* bl, int r, const float* __restrict__ cond,
                                                 int n, int c_begin, int c_end) {
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c * 8 + j;
      v[j] = k < n ? __ldg(cond + k) : 0.f;
    }
    store_chunk_x3(bh, bl, r, c, v);
  }
}

// Epilogue scheduling options (A/B builds: tools/build_variant.py x -DNFB_X3_OPT=<mask>).
#ifndef NFB_X3_OPT
#define NFB_X3_OPT 73
#endif
constexpr bool kOBias = (NFB_X3_OPT & 1) != 0;       // biases into registers before the wait for the accumulator
constexpr bool kOLdSplit = (NFB_X3_OPT & 2) != 0;    // second tcgen05.ld in flight under the first piece's arithmetic
constexpr bool kOPrefetch = (NFB_X3_OPT & 4) != 0;   // next tile's z / ray loads issued one or two steps early
constexpr bool kOCondEarly = (NFB_X3_OPT & 8) != 0;  // rgb condition written before the wait of its step
constexpr bool kOAdotLate = (NFB_X3_OPT & 16) != 0;  // alpha-head dot product after the layer's hand-off
constexpr bool kOEarlyNext = (NFB_X3_OPT & 32) != 0; // next tile's first input block encoded before the rgb head
constexpr bool kOBiasLdg = (NFB_X3_OPT & 64) != 0;   // biases through L1 (ld.global.nc) instead of the constant bank

// A step's biases: 18 KB over the steps of a level - more than the constant cache holds, so the indexed
// LDCs of the constant-bank copy miss (ncu: LDC + the FFMA2s waiting for them = 38 % of the epilogue's
// busy samples).  kOBiasLdg reads the same values from the global copy through L1.
__device__ __forceinline__ float4 ld_bias(const float4* p, int i) {
  if (kOBiasLdg) return __ldg(p + i);
  return p[i];
}

// One 32-column piece of a hidden layer's epilogue: + bias, activation, (alpha head
// dot product in fp32), split into fp16 hi / lo pairs.
// `inv_s` undoes the power-of-two weight scale of the layer (x3_weight_scale): acc * inv_s is
// exact, so fma(acc, inv_s, bias) rounds exactly like the unscaled acc + bias.
//
// The split.  hi = v with the fp32 mantissa truncated to fp16's 11 significant bits (one
// LOP3; the conversion to fp16 is then exact), lo = fp16(v - hi) (the subtraction is exact;
// lo carries the next 11 of the remaining 13 bits): v - (hi + lo) <= 2^-23 |v|, the same
// bound as a round-to-nearest split, for 5 instructions per two elements (2 LOP3, FADD2,
// 2 F2FP) instead of 8.  ReLU rides on the conversions (cvt.relu): for v < 0 both
// hi = trunc(v) and v - trunc(v) are <= 0 and convert to +0.  Below fp16's normal range
// (|v| < 2^-14) the conversion of hi rounds to the subnormal grid and lo does not see that
// error: an absolute 2^-25, far below the fp32 noise of a layer output.  Saturating
// conversions: |v| > 65504 does not become inf.
template <bool kRelu>
__device__ __forceinline__ void x3_piece_fast(float* v, const float4* __restrict__ bq4, float inv_s,
                                              uint32_t* hi16, uint32_t* lo16) {
  const uint64_t is2 = pack_f32x2(inv_s, inv_s);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 bq = kOBias ? bq4[j >> 2] : ld_bias(bq4, j >> 2);   // (kOBias: already in registers)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint64_t r, t, d;
      asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack_f32x2(v[j + 2 * h], v[j + 2 * h + 1])), "l"(is2),
          "l"(h == 0 ? pack_f32x2(bq.x, bq.y) : pack_f32x2(bq.z, bq.w)));
      // the pre-activation values stay in v[] (register renaming only; dead unless the alpha head reads them)
      asm("mov.b64 {%0, %1}, %2;" : "=f"(v[j + 2 * h]), "=f"(v[j + 2 * h + 1]) : "l"(r));
      asm("and.b64 %0, %1, 0xFFFFE000FFFFE000;" : "=l"(t) : "l"(r));
      asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(r), "l"(t));
      float t0, t1, d0, d1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
      if (kRelu) {
        asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi16[(j >> 1) + h]) : "f"(t1), "f"(t0));
        asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo16[(j >> 1) + h]) : "f"(d1), "f"(d0));
      } else {
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi16[(j >> 1) + h]) : "f"(t1), "f"(t0));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo16[(j >> 1) + h]) : "f"(d1), "f"(d0));
      }
    }
  }
}
// The alpha head (Dense(1) on the trunk output, one layer per level): an fp32 FMA chain over the
// activated values x3_piece left in v[], evaluated AFTER the layer's hand-off (the head's result is
// not needed before the rgb step) in the order the fused form used (columns ascending).
__device__ __forceinline__ void alpha_dot32(const float* v, const float4* __restrict__ aw4, bool relu, float& alpha) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 w = aw4[j >> 2];
    float t0 = v[j], t1 = v[j + 1], t2 = v[j + 2], t3 = v[j + 3];
    if (relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f); }
    alpha = fmaf(t0, w.x, alpha); alpha = fmaf(t1, w.y, alpha);
    alpha = fmaf(t2, w.z, alpha); alpha = fmaf(t3, w.w, alpha);
  }
}
__device__ __forceinline__ void x3_piece(float* v, const float4* __restrict__ bq4, float inv_s, bool relu,
                                         uint32_t* hi16, uint32_t* lo16) {
#ifdef NFB_X3_EXP_NOEPI
#pragma unroll
  for (int j = 0; j < 16; ++j) { hi16[j] = 0x3c003c00u; lo16[j] = 0u; }
  return;
#endif
  if (relu) x3_piece_fast<true>(v, bq4, inv_s, hi16, lo16);
  else x3_piece_fast<false>(v, bq4, inv_s, hi16, lo16);
}

// First chain of a unit: fence + 4 x (x_hi W_hi).  kTs: A from tensor memory (a = TMEM
// address, K steps of 8 columns) or from shared memory (a = descriptor, K steps of 32 bytes).
// Default (off with -DNFB_X3_NO_COLLECT): K-step-major order inside a unit so that x_hi[k] is read from tensor
// memory once for its two products (collector::a::fill, then ::lastuse):
//   for k: x_hi[k] W_hi[k] (fill), x_hi[k] W_lo[k] (lastuse), x_lo[k] W_hi[k]
#ifdef NFB_X3_NO_COLLECT
constexpr bool kCollect = false;
#else
constexpr bool kCollect = true;     // +0.5 % (three A/B rounds, profiles/r02_ab_x3_variants.txt): one TMEM read of x_hi less per K step
#endif
template <bool kTs>
__device__ __forceinline__ void issue_x3_head(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                              uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTs && kCollect) {
    asm volatile(
        "{\n\t"
        ".reg .pred pacc, pt;\n\t"
        ".reg .b32 h1;\n\t"
        ".reg .b64 bh1;\n\t"
        "tcgen05.fence::after_thread_sync;\n\t"
        "setp.ne.b32 pacc, %6, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "add.u32 h1, %1, 8;\n\t"
        "add.u64 bh1, %3, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], [%1], %3, %5, pacc;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], [%1], %4, %5, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %5, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], [h1], bh1, %5, pt;\n\t"
        "}"
        ::"r"(d), "r"((uint32_t)a_hi), "r"((uint32_t)a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
        : "memory");
  } else if constexpr (kTs) {
    asm volatile(
        "{\n\t"
        ".reg .pred pacc, pt;\n\t"
        ".reg .b32 a1, a2, a3;\n\t"
        ".reg .b64 b1, b2, b3;\n\t"
        "tcgen05.fence::after_thread_sync;\n\t"
        "setp.ne.b32 pacc, %4, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "add.u32 a1, %1, 8;\n\t add.u32 a2, %1, 16;\n\t add.u32 a3, %1, 24;\n\t"
        "add.u64 b1, %2, 2;\n\t add.u64 b2, %2, 4;\n\t add.u64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, pacc;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, pt;\n\t"
        "}"
        ::"r"(d), "r"((uint32_t)a_hi), "l"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    issue_half0(d, a_hi, b_hi, idesc, accumulate);
  }
  (void)a_lo; (void)b_lo;
}

// Second and third MMA chains of a unit plus everything that follows them:
//   4 x (x_lo W_hi), 4 x (x_hi W_lo) with the look-ahead probes of the next unit's
//   barriers in between, ONE commit "weight slot free" and the optional x_free /
//   accumulator commits.  Returns the probe bits: 1 = the next unit's weight slot has
//   landed, 2/4/8 = x_ready[0/1/2].
#define NFB_X3_TAIL_PROLOGUE \
      ".reg .pred pt, pw, px0, px1, px2, pd0, pd1, pd2, pcx, pca, pmc, pe, pe1, pe2;\n\t" \
      ".reg .b32 t0, t1, t2;\n\t" \
      "setp.eq.b32 pt, 0, 0;\n\t" \
      "setp.ne.b32 pmc, %17, 0;\n\t" \
      "setp.ne.b32 pe, %7, 0;\n\t" \
      "and.pred pe2, pe, pmc;\n\t" \
      "not.pred pe1, pmc;\n\t" \
      "and.pred pe1, pe1, pe;\n\t" \
      "setp.ne.b32 pd0, %12, 0;\n\t" \
      "setp.ne.b32 pd1, %13, 0;\n\t" \
      "setp.ne.b32 pd2, %14, 0;\n\t" \
      "setp.ne.b32 pcx, %8, 0;\n\t" \
      "setp.ne.b32 pca, %9, 0;\n\t" \
      "setp.eq.b32 px0, 1, 0;\n\t" \
      "setp.eq.b32 px1, 1, 0;\n\t" \
      "setp.eq.b32 px2, 1, 0;\n\t" \
      "add.u64 bh1, %4, 2;\n\t add.u64 bh2, %4, 4;\n\t add.u64 bh3, %4, 6;\n\t" \
      "add.u64 bl1, %5, 2;\n\t add.u64 bl2, %5, 4;\n\t add.u64 bl3, %5, 6;\n\t"
#define NFB_X3_TAIL_PROBES \
      "mbarrier.test_wait.parity.shared::cta.b64 pw, [%10], %11;\n\t" \
      "@pd0 mbarrier.test_wait.parity.shared::cta.b64 px0, [%12], %15;\n\t" \
      "@pd1 mbarrier.test_wait.parity.shared::cta.b64 px1, [%13], %15;\n\t" \
      "@pd2 mbarrier.test_wait.parity.shared::cta.b64 px2, [%14], %15;\n\t"
#define NFB_X3_TAIL_EPILOGUE \
      /* slot release: every second unit (a PAIR of slots per commit: a commit costs ~120 cycles of issue) */ \
      "@pe1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t" \
      /* cluster build: the weight slot is shared (multicast copies): release it in every CTA of the cluster */ \
      "@pe2 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], %16;\n\t" \
      "@pcx tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t" \
      "@pca tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t" \
      "selp.u32 %0, 1, 0, pw;\n\t" \
      "selp.u32 t0, 2, 0, px0;\n\t" \
      "selp.u32 t1, 4, 0, px1;\n\t" \
      "selp.u32 t2, 8, 0, px2;\n\t" \
      "or.b32 %0, %0, t0;\n\t" \
      "or.b32 %0, %0, t1;\n\t" \
      "or.b32 %0, %0, t2;\n\t"
template <int kCl, bool kTs>
__device__ __forceinline__ uint32_t issue_x3_tail(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                                  uint64_t b_lo, uint32_t idesc, uint32_t bar_empty,
                                                  uint32_t bar_xfree, uint32_t bar_acc, uint32_t probe_w,
                                                  uint32_t par_w, uint32_t probe_x0, uint32_t probe_x1,
                                                  uint32_t probe_x2, uint32_t par_x) {
  uint32_t out;
  if constexpr (kTs && kCollect) {
    asm volatile(
        "{\n\t"
        ".reg .b32 l1, l2, l3, h1, h2, h3;\n\t"
        ".reg .b64 bh1, bh2, bh3, bl1, bl2, bl3;\n\t"
        NFB_X3_TAIL_PROLOGUE
        "add.u32 l1, %3, 8;\n\t add.u32 l2, %3, 16;\n\t add.u32 l3, %3, 24;\n\t"
        "add.u32 h1, %2, 8;\n\t add.u32 h2, %2, 16;\n\t add.u32 h3, %2, 24;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%1], [h1], bl1, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [l1], bh1, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%1], [h2], bh2, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%1], [h2], bl2, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [l2], bh2, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%1], [h3], bh3, %6, pt;\n\t"
        NFB_X3_TAIL_PROBES
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%1], [h3], bl3, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [l3], bh3, %6, pt;\n\t"
        NFB_X3_TAIL_EPILOGUE
        "}"
        : "=r"(out)
        : "r"(d), "r"((uint32_t)a_hi), "r"((uint32_t)a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(bar_empty),
          "r"(bar_xfree), "r"(bar_acc), "r"(probe_w), "r"(par_w), "r"(probe_x0), "r"(probe_x1),
          "r"(probe_x2), "r"(par_x), "h"((uint16_t)((1u << kCl) - 1u)), "r"(kCl > 1 ? 1u : 0u)
        : "memory");
  } else if constexpr (kTs) {
    asm volatile(
        "{\n\t"
        ".reg .b32 l1, l2, l3, h1, h2, h3;\n\t"
        ".reg .b64 bh1, bh2, bh3, bl1, bl2, bl3;\n\t"
        NFB_X3_TAIL_PROLOGUE
        "add.u32 l1, %3, 8;\n\t add.u32 l2, %3, 16;\n\t add.u32 l3, %3, 24;\n\t"
        "add.u32 h1, %2, 8;\n\t add.u32 h2, %2, 16;\n\t add.u32 h3, %2, 24;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [%3], %4, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [l1], bh1, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [l2], bh2, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [l3], bh3, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %5, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [h1], bl1, %6, pt;\n\t"
        NFB_X3_TAIL_PROBES
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [h2], bl2, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [h3], bl3, %6, pt;\n\t"
        NFB_X3_TAIL_EPILOGUE
        "}"
        : "=r"(out)
        : "r"(d), "r"((uint32_t)a_hi), "r"((uint32_t)a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(bar_empty),
          "r"(bar_xfree), "r"(bar_acc), "r"(probe_w), "r"(par_w), "r"(probe_x0), "r"(probe_x1),
          "r"(probe_x2), "r"(par_x), "h"((uint16_t)((1u << kCl) - 1u)), "r"(kCl > 1 ? 1u : 0u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .b64 l1, l2, l3, h1, h2, h3, bh1, bh2, bh3, bl1, bl2, bl3;\n\t"
        NFB_X3_TAIL_PROLOGUE
        "add.u64 l1, %3, 2;\n\t add.u64 l2, %3, 4;\n\t add.u64 l3, %3, 6;\n\t"
        "add.u64 h1, %2, 2;\n\t add.u64 h2, %2, 4;\n\t add.u64 h3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %4, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], l1, bh1, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], l2, bh2, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], l3, bh3, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %5, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], h1, bl1, %6, pt;\n\t"
        NFB_X3_TAIL_PROBES
        "tcgen05.mma.cta_group::1.kind::f16 [%1], h2, bl2, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], h3, bl3, %6, pt;\n\t"
        NFB_X3_TAIL_EPILOGUE
        "}"
        : "=r"(out)
        : "r"(d), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(bar_empty),
          "r"(bar_xfree), "r"(bar_acc), "r"(probe_w), "r"(par_w), "r"(probe_x0), "r"(probe_x1),
          "r"(probe_x2), "r"(par_x), "h"((uint16_t)((1u << kCl) - 1u)), "r"(kCl > 1 ? 1u : 0u)
        : "memory");
  }
  return out;
}
#undef NFB_X3_TAIL_PROLOGUE
#undef NFB_X3_TAIL_PROBES
#undef NFB_X3_TAIL_EPILOGUE

// The 32 biases of a piece, fetched from the constant bank BEFORE the wait for the accumulator (the
// empty asm keeps the loads there: otherwise they are sunk to their uses and their latency lands
// on the hand-off chain accumulator -> epilogue -> next layer's first MMA).
__device__ __forceinline__ const float4* bias_preload(const float4* __restrict__ src, float4* dst) {
  if (!kOBias) return src;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dst[i] = ld_bias(src, i);
    asm volatile("" ::"f"(dst[i].x), "f"(dst[i].y), "f"(dst[i].z), "f"(dst[i].w));
  }
  return dst;
}

// One 32-column piece of a layer's output (16 packed pairs per image) -> this thread's TMEM lane.
// (NFB_X3_EXP_*: timing experiments of developer builds, garbage results.)
__device__ __forceinline__ void tst_piece(uint32_t t_lane, int col, const uint32_t* hi16, const uint32_t* lo16) {
#if !defined(NFB_X3_EXP_NOTMEM) && !defined(NFB_X3_EXP_NOEPI)
  tmem_st16(t_lane + kTmAHi + (uint32_t)(col >> 1), hi16);
  tmem_st16(t_lane + kTmALo + (uint32_t)(col >> 1), lo16);
#endif
}
__device__ __forceinline__ void x3_ld32(uint32_t taddr, float* v) {
#if !defined(NFB_X3_EXP_NOTMEM) && !defined(NFB_X3_EXP_NOEPI)
  tmem_ld32(taddr, v);
#else
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 1.f;
#endif
}

// Row state owned by the two epilogue threads of a row for the lifetime of a tile.
struct X3Row {
  float x[3];        // current (possibly warped) sample point
  long long m;       // global row (clamped to a valid row)
  long long ray;
  bool valid;
  float alpha;       // this thread's part of the alpha-head dot product
  float z, dist;     // fused composite: z of the sample, (z_next - z) * |d| (or the last-sample constant)
  bool last;         // last sample of its ray
};

// kCl > 1: launched as clusters of kCl CTAs that SHARE the weight stream: rank i fetches part i of
// every slot with a multicast bulk copy into all the CTAs' rings.
// Every SM of the chip streams the same 5.6 MB of weights in near lockstep, which makes the L2
// slices that hold the current unit the bottleneck (148 readers per line: a 32 KB slot took ~1,500
// cycles to arrive); sharing halves those reads.  Otherwise the two CTAs are independent (own
// tiles, own MMAs, own TMEM): the only coupling is that a slot is refilled once BOTH issuers have
// released it (empty barriers count two multicast commits).
template <int kCl>
__global__ void __launch_bounds__(kX3Threads, 1)
field_x3_kernel(const __grid_constant__ TcProgram prog, const __grid_constant__ X3Consts cst,
                const FieldArgs args, const uint8_t* __restrict__ wpack, int num_tiles,
                const float* __restrict__ aux) {
  constexpr int kEpiWarps = 8, kMmaWarp = 8, kProdWarp = 9;
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* base = raw;
  if ((smem_u32(base) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("nfb: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* inh = base + kInhOff;               // input block, hi
  uint8_t* inl = base + kInlOff;
  uint8_t* stages = base + kStageOff;
  float* alpha_part = reinterpret_cast<float*>(base + kPartOff);
  X3Bars* bars = reinterpret_cast<X3Bars*>(base + kBarOff);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == kMmaWarp * 32) {
    for (int i = 0; i < kX3Slots; ++i) {
      mbar_init(&bars->full[i], 1);
      if ((i & 1) == 0) mbar_init(&bars->empty[i >> 1], kCl);
    }
    mbar_init(&bars->acc_ready[0], 1); mbar_init(&bars->acc_ready[1], 1);
    mbar_init(&bars->x_free, 1);
    for (int k = 0; k < 3; ++k) mbar_init(&bars->x_ready[k], kXrCount);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&bars->tmem_slot, kTmCols);
  tc_fence_before();
  if constexpr (kCl > 1) cluster_sync_all();    // the peers' barriers exist before any multicast lands
  else __syncthreads();
  tc_fence_after();
  const uint32_t rank = kCl > 1 ? cluster_ctarank() : 0u;
  const uint32_t tmem_base = bars->tmem_slot;
  const bool do_warp = args.use_warp && prog.warp_type != 0;
  int first_step = 0;
  if (!do_warp) {
    while (first_step < prog.n_steps && prog.steps[first_step].epi != kEpiWarpHeads) ++first_step;
    first_step = (first_step < prog.n_steps) ? first_step + 1 : 0;   // skip the warp net
  }
  int last_step = prog.n_steps - 1;
  if (args.warp_only) {
    last_step = 0;
    while (prog.steps[last_step].epi != kEpiWarpHeads) ++last_step;
  }
  constexpr int kCtlRegs = 40, kEpiRegs = 232;     // 128 x 40 + 256 x 232 = 384 x 168
  // Tiles of this CTA.  Default: blockIdx.x, + gridDim.x, ...  With the fused composite a CTA
  // takes whole rays - the S / 128 tiles of ray blockIdx.x (+ gridDim.x, ...) back to back - so
  // that transmittance and the partial sums are carried in registers from tile to tile.
  const bool fuse = args.ray_out != nullptr && !args.warp_only;
  const int tpr = fuse ? args.samples_per_ray / kTileRows : 1;          // tiles per group
  const int groups = num_tiles / tpr;
  // (the CTAs of a cluster run the same number of units - the count of the first member: the ring couples
  //  them; the odd member's surplus tile lies beyond the end, is computed on clamped rows and never stored)
  const int bid_n = (int)blockIdx.x & ~(kCl - 1);
  const int my_groups = bid_n < groups ? (groups - bid_n + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n_my = my_groups * tpr;
  auto tile_of = [&](int i) { return ((int)blockIdx.x + (i / tpr) * (int)gridDim.x) * tpr + (i % tpr); };

  if (warp == kProdWarp) {
    // ===================== weight producer =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
    uint32_t sg = 0, ph = 0, dead = 0;
    Tracer tr(args, lane == 0 ? 3 : -1);
    for (int ti = 0; ti < n_my; ++ti) {
      for (int si = first_step; si <= last_step; ++si) {
        const TcStep& st = prog.steps[si];
        const uint32_t bytes = 2u * (uint32_t)st.chunk_n * kRowBytes;   // [W_hi | W_lo] of one unit, contiguous
        const uint8_t* src = wpack + st.w_off;
        const int n = st.n_chunks * st.nkb;               // [chunk][kb]
        for (int u = 0; u < n; ++u) {
          uint8_t* dst = stages + sg * kX3SlotBytes;
          const uint8_t* from = src + (size_t)u * bytes;
          if ((sg & 1) == 0) mbar_wait(&bars->empty[sg >> 1], ph ^ 1, dead);
          tr.ev(si, u);
#ifdef NFB_X3_EXP_NOFILL
          // timing experiment (garbage results): only the first pass over the ring is copied
          if (ti > 0 || si > first_step + 1) { if (elect_one()) mbar_arrive(&bars->full[sg]); __syncwarp();
            if (++sg == kX3Slots) { sg = 0; ph ^= 1; } continue; }
#endif
          if (elect_one()) {
            mbar_arrive_expect_tx(&bars->full[sg], bytes);
            if constexpr (kCl > 1) {
              // every rank fetches its 1 / kCl of the slot - to all CTAs of the cluster
              const uint32_t part = bytes / kCl;
              bulk_g2s_multicast(dst + rank * part, from + rank * part, part, &bars->full[sg], (uint16_t)((1u << kCl) - 1u));
            } else {
              bulk_g2s(dst, from, bytes, &bars->full[sg]);
            }
          }
          __syncwarp();
          if (++sg == kX3Slots) { sg = 0; ph ^= 1; }
        }
      }
    }
    if (lane == 0) tr.finish(args, 3);
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
    if (elect_one()) {
      Tracer tr(args, 0);
      const uint64_t desc_hi = make_smem_desc(0) & 0xFFFFFFFF00000000ull;
      const uint32_t lo_base = ((smem_u32(base) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t st_lo = ((smem_u32(stages) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_full = smem_u32(&bars->full[0]), b_empty = smem_u32(&bars->empty[0]);
      const uint32_t b_acc0 = smem_u32(&bars->acc_ready[0]), b_acc1 = smem_u32(&bars->acc_ready[1]);
      const uint32_t b_xfree = smem_u32(&bars->x_free);
      const uint32_t b_x0 = smem_u32(&bars->x_ready[0]), b_x1 = smem_u32(&bars->x_ready[1]);
      const uint32_t b_x2 = smem_u32(&bars->x_ready[2]);
      const int u_begin = prog.unit_begin[first_step], u_end = prog.unit_begin[last_step + 1];
      const int n_u = u_end - u_begin;
      const uint4* utab = reinterpret_cast<const uint4*>(prog.units);   // 2 x uint4 per unit
      uint32_t sg = 0, wph = 0, xr = 0, ready = 0, dead = 0;
      uint4 c0 = utab[2 * u_begin], c1 = utab[2 * u_begin + 1];
      int un = u_begin + (1 % n_u);
      uint4 n0 = utab[2 * un], n1 = utab[2 * un + 1];
      uint32_t d0 = tmem_base + c0.z;
      uint64_t bd = desc_hi | (uint64_t)st_lo;
      // A operand of a K-block: TMEM address (activations) or shared-memory descriptor (input block)
      auto a_op = [&](uint32_t v) -> uint64_t {
        return (v & kUnitATmem) ? (uint64_t)(tmem_base + (v & 0xffffu)) : (desc_hi | (uint64_t)(lo_base + v));
      };
      uint64_t ad0 = a_op(c0.x);
      for (int ti = 0; ti < n_my; ++ti) {
        for (int u = u_begin; u < u_end; ++u) {
          const uint32_t flags = c1.x, need = c1.y;
          tr.ev(c1.w, 64 + (need & ~ready));               // what (if anything) this unit must wait for
          if ((ready & need) != need) {                    // slow path: something is not there yet
            if ((need & 2) && !(ready & 2)) mbar_wait_issuer(&bars->x_ready[0], xr & 1, dead);
            if ((need & 4) && !(ready & 4)) mbar_wait_issuer(&bars->x_ready[1], xr & 1, dead);
            if ((need & 8) && !(ready & 8)) mbar_wait_issuer(&bars->x_ready[2], xr & 1, dead);
            if (!(ready & 1)) mbar_wait_issuer(&bars->full[sg], wph, dead);
          }
          const bool ts = (c0.x & kUnitATmem) != 0;
          const uint64_t a_lo = a_op(c0.y);
          const uint64_t b_lo = bd + (uint64_t)(c1.z >> 16);    // W_lo follows W_hi inside the slot
          if (ts) issue_x3_head<true>(d0, ad0, a_lo, bd, b_lo, c0.w, flags & kUAccum);     // x_hi W_hi
          else issue_x3_head<false>(d0, ad0, a_lo, bd, b_lo, c0.w, flags & kUAccum);
          // ---- bookkeeping while those MMAs execute ----
          if (++un >= u_end) un -= n_u;
          const uint4 f0 = utab[2 * un], f1 = utab[2 * un + 1];          // table entry two units ahead
          const uint32_t nsg = sg + 1 == kX3Slots ? 0u : sg + 1;
          const uint32_t nwph = wph ^ (nsg == 0 ? 1u : 0u);
          const uint32_t nxr = xr + ((flags & kUStepEnd) ? 1u : 0u);
          const uint32_t px0 = (c1.z & 2) ? b_x0 : 0u, px1 = (c1.z & 4) ? b_x1 : 0u;
          const uint32_t px2 = (c1.z & 8) ? b_x2 : 0u;
          const uint32_t d_cur = d0, idesc = c0.w, bar_e = (sg & 1) ? b_empty + (sg >> 1) * 8 : 0u;
          const uint64_t bd_cur = bd, a_hi = ad0;
          d0 = tmem_base + n0.z;
          bd = desc_hi | (uint64_t)(st_lo + nsg * (kX3SlotBytes >> 4));
          ad0 = a_op(n0.x);
          // W_lo follows W_hi inside the slot: chunk_n rows x 128 B further (c1.z bits 16..)
          const uint32_t bar_x = (flags & kUCommitXFree) ? b_xfree : 0u;
          const uint32_t bar_a = (flags & kUCommitAcc0) ? b_acc0 : ((flags & kUCommitAcc1) ? b_acc1 : 0u);
          if (ts)
            ready = issue_x3_tail<kCl, true>(d_cur, a_hi, a_lo, bd_cur, b_lo, idesc, bar_e, bar_x, bar_a,
                                               b_full + nsg * 8, nwph, px0, px1, px2, nxr & 1);
          else
            ready = issue_x3_tail<kCl, false>(d_cur, a_hi, a_lo, bd_cur, b_lo, idesc, bar_e, bar_x, bar_a,
                                                b_full + nsg * 8, nwph, px0, px1, px2, nxr & 1);
          if (flags & (kUCommitAcc0 | kUCommitAcc1)) tr.ev(c1.w, (flags & kUCommitAcc0) ? 1 : 2);
          if (flags & kUWaitX0) tr.ev(c1.w, 0);
          sg = nsg; wph = nwph; xr = nxr;
          c0 = n0; c1 = n1; n0 = f0; n1 = f1;
        }
      }
      tr.finish(args, 0);
    }
    __syncwarp();
  } else if (warp >= kEpiWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));   // idle warps of the control group
  } else {
    // ===================== epilogue: two threads per row =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const int hs = warp >> 2;                          // which column half of a chunk
    const int qd = warp & 3;                           // TMEM lane quarter
    const int r = qd * 32 + lane;                      // row within the tile
    const uint32_t t_lane = tmem_base + (((uint32_t)qd * 32) << 16);
    const int cb = hs * 4, ce = cb + 4;                // input-block chunks this thread writes
    Tracer tr(args, (lane == 0 && qd == 0) ? 1 + hs : -1);
    uint32_t n_acc0 = 0, n_acc1 = 0, n_free = 0, dead = 0;
    X3Row row;
    const int S = args.samples_per_ray;
    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kX3EpiThreads) : "memory"); };
    auto arrive_all = [&]() {
      fence_proxy_async();
      tc_fence_before();
#ifdef NFB_X3_WARP_ARRIVE
      __syncwarp();
      if (lane == 0) { mbar_arrive(&bars->x_ready[0]); mbar_arrive(&bars->x_ready[1]); mbar_arrive(&bars->x_ready[2]); }
#else
      xr_arrive(&bars->x_ready[0]);
      xr_arrive(&bars->x_ready[1]);
      xr_arrive(&bars->x_ready[2]);
#endif
    };
    // Sample point of this thread's row in tile `tile`, and the first input block
    // (model_utils.py:72-73; warping.py:325-326 / models.py:270).
    // The global loads of a tile's rows (z, ray origin / direction), issued EARLY - before the wait
    // for the previous tile's last accumulator - so that their L2 latency is off the serial section
    // between two tiles.
    struct TilePref { float z, zn, org[3], dir[3]; };
    auto tile_prefetch = [&](int tile, TilePref& pf) {
      long long m = (long long)tile * kTileRows + r;
      if (m >= args.num_rows) m = args.num_rows - 1;
      const long long ray = m / S;
      pf.z = args.z_vals ? __ldg(args.z_vals + m) : 0.f;
      const bool last = m + 1 == (ray + 1) * S;
      pf.zn = (fuse && !last && args.z_vals) ? __ldg(args.z_vals + m + 1) : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        pf.dir[c] = __ldg(args.directions + ray * 3 + c);
        pf.org[c] = __ldg(args.origins + ray * 3 + c);
      }
    };
    // Row state of tile `tile` (model_utils.py:72-73) ...
    auto tile_row = [&](int tile, const TilePref& pf, X3Row& rw) {
      long long m = (long long)tile * kTileRows + r;
      rw.valid = m < args.num_rows;
      if (!rw.valid) m = args.num_rows - 1;
      rw.m = m;
      rw.ray = m / S;
      const float z = pf.z;
      float dir[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dir[c] = pf.dir[c];
        rw.x[c] = pf.org[c] + z * dir[c];
      }
      if (fuse) {
        // dists of volumetric_rendering (model_utils.py:98-104)
        rw.z = z;
        rw.last = m + 1 == (rw.ray + 1) * S;
        const float dnorm = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        const float d = rw.last ? (args.sample_at_infinity ? 1e10f : 1e-19f) : (pf.zn - z);
        rw.dist = d * dnorm;
      }
      if (!do_warp && args.warped && rw.valid && hs == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) args.warped[m * 3 + c] = rw.x[c];
      }
      rw.alpha = hs == 0 ? cst.alpha_b : 0.f;
    };
    // ... and this thread's half of its first input block (warping.py:325-326 / models.py:270) as
    // packed fp16 hi / lo pairs in registers.
    auto tile_encode = [&](const X3Row& rw, uint4* qh, uint4* ql) {
      const float* cond = args.cond + rw.ray * prog.cond_stride;
      if (do_warp) {
        if (hs == 0) posenc_pack_x3<0>(qh, ql, rw.x, prog.Fw, args.window, cond, prog.G);
        else posenc_pack_x3<1>(qh, ql, rw.x, prog.Fw, args.window, cond, prog.G);
      } else {
        if (hs == 0) posenc_pack_x3<0>(qh, ql, rw.x, prog.Fp, nullptr, nullptr, 0);
        else posenc_pack_x3<1>(qh, ql, rw.x, prog.Fp, nullptr, nullptr, 0);
      }
    };
    auto begin_tile = [&](int tile, const TilePref& pf) {
      uint4 qh[4], ql[4];
      tile_row(tile, pf, row);
      tile_encode(row, qh, ql);
      store_packed_x3(inh, inl, r, hs, qh, ql);
    };

    // fused composite: running state of the ray this CTA is on (replicated in every thread)
    float* scan_s = reinterpret_cast<float*>(base + kScanOff);
    float c_T = 1.f, c_cw = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f, a_w = 0.f, a_wnl = 0.f, a_med = 0.f;
    TilePref pf;
    if (n_my > 0) {
      tile_prefetch(tile_of(0), pf);
      begin_tile(tile_of(0), pf);
      arrive_all();
    }
    // The serial section between two tiles.  The next tile's rows are fetched two steps before the
    // end of a tile and its first input block is ENCODED (fp64 sines, ~3.5 K cycles) in the idle time
    // before the accumulator of the step in front of the rgb head arrives; the registers are stored
    // once that step's MMAs - the last readers of the input block - are complete.  The rgb-head
    // epilogue then releases the issuer as soon as it has drained its accumulator, and the volumetric
    // rendering of this tile runs under the first MMAs of the next one.  (Needs an rgb head that does
    // not read the input block; otherwise the encoding is written after the rgb head, as before.)
    bool early_ok = !args.warp_only && prog.steps[last_step].epi == kEpiRgbOut && last_step - 1 >= first_step &&
                    prog.steps[last_step - 1].epi == kEpiHidden;
    for (int kb = 0; kb < prog.steps[last_step].nkb; ++kb)
      if (prog.steps[last_step].src[kb] == kSrcIn) early_ok = false;
    if (!kOEarlyNext) early_ok = false;
    // (without kOPrefetch the loads are issued right in front of begin_tile: pf_step = -1)
    const int pf_step = early_ok ? (last_step - 2 >= first_step ? last_step - 2 : last_step - 1)
                                 : (kOPrefetch ? last_step : -1);
    X3Row nrow;
    uint4 nqh[4], nql[4];
    bool next_stored = false;
    for (int ti = 0; ti < n_my; ++ti) {
      const int tile = tile_of(ti);
      const bool has_next = ti + 1 < n_my;
      for (int si = first_step; si <= last_step; ++si) {
        const TcStep& st = prog.steps[si];
        const float4* bias4 = kOBiasLdg ? reinterpret_cast<const float4*>(aux + st.b_off)
                                        : cst.b4 + si * 64;   // this step's 256 biases
        const float inv_s = cst.inv_scale[si];          // undoes the step's power-of-two weight scale
        if (has_next && si == pf_step && !args.warp_only) tile_prefetch(tile_of(ti + 1), pf);
        const bool encode_here = early_ok && has_next && si == last_step - 1;
        if (encode_here) {
          tile_row(tile_of(ti + 1), pf, nrow);
          tile_encode(nrow, nqh, nql);
        }
        if (st.epi == kEpiHidden && st.n_chunks == 1) {
          // ---- 128-wide layer, one N = 128 chunk: every MMA of the layer is complete, so the
          //      output overwrites the input in place; two 64-column instalments (x_ready[0] /
          //      x_ready[2]) so that the next layer's first K-block starts under the second ----
          const bool relu = st.relu != 0;
          uint32_t ph[16], pl[16];
          float va[32], vb[32];
          const int ca = hs * 32, cb2 = 64 + hs * 32;
          float4 bAr[8], bBr[8];
          const float4* bA = bias_preload(bias4 + (ca >> 2), bAr);
          const float4* bB = bias_preload(bias4 + (cb2 >> 2), bBr);
          mbar_wait(&bars->acc_ready[0], n_acc0++ & 1, dead);
          tc_fence_after();
          tr.ev(si, 0);
          x3_ld32(t_lane + ca, va);
          if (kOLdSplit) {
            tmem_ld_wait();
            x3_ld32(t_lane + cb2, vb);                  // in flight under the first piece's arithmetic
            x3_piece(va, bA, inv_s, relu, ph, pl);
            tmem_ld_wait();
            tc_fence_before();
            xr_arrive(&bars->x_ready[1]);               // the accumulator may be overwritten
          } else {
            x3_ld32(t_lane + cb2, vb);
            tmem_ld_wait();
            tc_fence_before();
            xr_arrive(&bars->x_ready[1]);               // the accumulator may be overwritten
            x3_piece(va, bA, inv_s, relu, ph, pl);
          }
          tst_piece(t_lane, ca, ph, pl);
          tmem_st_wait();
          tc_fence_before();
          xr_arrive(&bars->x_ready[0]);
          tr.ev(si, 3);
          x3_piece(vb, bB, inv_s, relu, ph, pl);
          tst_piece(t_lane, cb2, ph, pl);
          tmem_st_wait();
          tc_fence_before();
          xr_arrive(&bars->x_ready[2]);
          tr.ev(si, 5);
        } else if (st.epi == kEpiHidden) {
          // ---- 256-wide layer, two N = 128 chunks; each thread owns 64 columns of a chunk ----
          const bool relu = st.relu != 0, adot = st.alpha_dot != 0;
          uint32_t ph[32], pl[32];
          float va[32], vb[32];
          const int col0 = hs * 64;
          const int col1 = st.chunk_n + col0;
          float4 bAr[8], bBr[8];
          const float4* bA = bias_preload(bias4 + (col0 >> 2), bAr);
          const float4* bB = bias_preload(bias4 + (col0 >> 2) + 8, bBr);
          // The rgb condition goes into the input block (the first K-block of the next step).  Every
          // earlier reader of the block belongs to a step whose last accumulator this thread has already
          // waited for, so it may be written BEFORE the wait (kOCondEarly) unless this very step reads the block.
          const bool cond_early = kOCondEarly && st.write_cond && st.src[0] != kSrcIn;
          if (cond_early) {
            cond_to_block_x3(inh, inl, r, args.cond + row.ray * prog.cond_stride + prog.G, prog.rc, cb, ce);
            fence_proxy_async();
          }
          // ---- chunk 0: results wait in registers until the MMAs of chunk 1 no
          //      longer read the blocks they overwrite ----
          mbar_wait(&bars->acc_ready[0], n_acc0++ & 1, dead);
          tc_fence_after();
          tr.ev(si, 0);
          x3_ld32(t_lane + col0, va);
          if (kOLdSplit) {
            tmem_ld_wait();
            x3_ld32(t_lane + col0 + 32, vb);            // in flight under the first piece's arithmetic
            x3_piece(va, bA, inv_s, relu, ph, pl);
            tmem_ld_wait();
          } else {
            x3_ld32(t_lane + col0 + 32, vb);
            tmem_ld_wait();
            x3_piece(va, bA, inv_s, relu, ph, pl);
          }
          if (adot && !kOAdotLate) alpha_dot32(va, cst.alpha4 + (col0 >> 2), relu, row.alpha);
          x3_piece(vb, bB, inv_s, relu, ph + 16, pl + 16);
          if (adot && !kOAdotLate) alpha_dot32(vb, cst.alpha4 + (col0 >> 2) + 8, relu, row.alpha);
          tr.ev(si, 1);
          if (st.kb_free != -2) mbar_wait(&bars->x_free, n_free++ & 1, dead);    // (-2: implied by acc_ready[0])
          tr.ev(si, 2);
          tst_piece(t_lane, col0, ph, pl);
          tst_piece(t_lane, col0 + 32, ph + 16, pl + 16);
          if (st.write_cond && !cond_early)
            cond_to_block_x3(inh, inl, r, args.cond + row.ray * prog.cond_stride + prog.G, prog.rc, cb, ce);
          tmem_st_wait();                               // the TMEM stores have completed ...
          if (st.write_cond && !cond_early) fence_proxy_async();   // ... and the input-block stores are visible to the MMAs
          tc_fence_before();
          xr_arrive(&bars->x_ready[0]);
          tr.ev(si, 3);
          if (adot && kOAdotLate) {                     // alpha head: after the hand-off
            alpha_dot32(va, cst.alpha4 + (col0 >> 2), relu, row.alpha);
            alpha_dot32(vb, cst.alpha4 + (col0 >> 2) + 8, relu, row.alpha);
          }
          // ---- chunk 1: every MMA of the layer is complete, store directly ----
          bA = bias_preload(bias4 + (col1 >> 2), bAr);
          bB = bias_preload(bias4 + (col1 >> 2) + 8, bBr);
          mbar_wait(&bars->acc_ready[1], n_acc1++ & 1, dead);
          tc_fence_after();
          tr.ev(si, 4);
          x3_ld32(t_lane + col1, va);
          if (kOLdSplit) {
            tmem_ld_wait();
            x3_ld32(t_lane + col1 + 32, vb);
            x3_piece(va, bA, inv_s, relu, ph, pl);
            tmem_ld_wait();
            tc_fence_before();
            xr_arrive(&bars->x_ready[1]);             // this accumulator may be overwritten (next step's chunk 1)
          } else {
            x3_ld32(t_lane + col1 + 32, vb);
            tmem_ld_wait();
            tc_fence_before();
            xr_arrive(&bars->x_ready[1]);
            x3_piece(va, bA, inv_s, relu, ph, pl);
          }
          if (adot && !kOAdotLate) alpha_dot32(va, cst.alpha4 + (col1 >> 2), relu, row.alpha);
          tst_piece(t_lane, col1, ph, pl);
          x3_piece(vb, bB, inv_s, relu, ph, pl);
          if (adot && !kOAdotLate) alpha_dot32(vb, cst.alpha4 + (col1 >> 2) + 8, relu, row.alpha);
          tst_piece(t_lane, col1 + 32, ph, pl);
          tmem_st_wait();
          tc_fence_before();
          xr_arrive(&bars->x_ready[2]);
          tr.ev(si, 5);
          if (adot) {
            if (kOAdotLate) {
              alpha_dot32(va, cst.alpha4 + (col1 >> 2), relu, row.alpha);
              alpha_dot32(vb, cst.alpha4 + (col1 >> 2) + 8, relu, row.alpha);
            }
            if (hs == 1) alpha_part[r] = row.alpha;     // read by the row's first thread at the rgb step
          }
        } else {
          // ---- heads: N = 16 accumulator columns, one chunk (both threads of a row
          //      do the scalar work; they split the input-block chunks) ----
          float v[16];
          mbar_wait(&bars->acc_ready[0], n_acc0++ & 1, dead);
          tc_fence_after();
          tr.ev(si, 0);
          tmem_ld16(t_lane, v);
          tmem_ld_wait();
          // rgb head with the next tile's input block already in place: the issuer may go on
          if (st.epi == kEpiRgbOut && next_stored) arrive_all();
#pragma unroll
          for (int j = 0; j < 12; j += 4) {
            const float4 bq = ld_bias(bias4, j >> 2);
            v[j] = fmaf(v[j], inv_s, bq.x); v[j + 1] = fmaf(v[j + 1], inv_s, bq.y);
            v[j + 2] = fmaf(v[j + 2], inv_s, bq.z); v[j + 3] = fmaf(v[j + 3], inv_s, bq.w);
          }
          if (st.epi == kEpiWarpHeads) {
            float y[3];
            if (prog.warp_type == 2) {
              se3_apply(v, row.x, y, prog.warp_pivot ? v + 6 : nullptr,
                        prog.warp_trans ? v + (prog.warp_pivot ? 9 : 6) : nullptr);
            } else {
#pragma unroll
              for (int c = 0; c < 3; ++c) y[c] = row.x[c] + v[c];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) row.x[c] = y[c];
            if (args.warped && row.valid && hs == 0) {
#pragma unroll
              for (int c = 0; c < 3; ++c) args.warped[row.m * 3 + c] = y[c];
            }
            if (args.warp_only) {
              if (has_next) { tile_prefetch(tile_of(ti + 1), pf); begin_tile(tile_of(ti + 1), pf); }
            } else {
              if (hs == 0) posenc_block_x3<0>(inh, inl, r, row.x, prog.Fp, nullptr, nullptr, 0);
              else posenc_block_x3<1>(inh, inl, r, row.x, prog.Fp, nullptr, nullptr, 0);
            }
            arrive_all();
          } else {
            epi_sync();                                 // the partner's alpha partial is visible
            // (the next write of alpha_part[r] is ordered behind this read by the barrier
            // chain of the next tile's layers: x_ready <- this thread, acc_ready -> writer)
            float4 o;
            if (hs == 0) {
              o.x = sigmoidf(v[0]); o.y = sigmoidf(v[1]); o.z = sigmoidf(v[2]);
              o.w = apply_act(row.alpha + alpha_part[r], prog.sigma_act);
              if (row.valid && args.samples) reinterpret_cast<float4*>(args.samples)[row.m] = o;
            }
            if (fuse && hs == 0) {
              // ---- volumetric_rendering (model_utils.py:104-136) + median depth (:231-239, 262-263)
              //      over the 128 samples of this tile; warps 0-3, one sample per thread ----
              auto bar128 = [&]() { asm volatile("bar.sync 2, 128;" ::: "memory"); };
              if ((tile % tpr) == 0) { c_T = 1.f; c_cw = 0.f; a_r = a_g = a_b = a_d = a_w = a_wnl = a_med = 0.f; }
              // alpha = 1 - exp(-sigma * dist) as -expm1(-x) (see composite_kernel)
              const float al = -expm1f(-o.w * row.dist);
              const float tf = 1.0f - al + 1e-10f;
              float P = tf;                                    // inclusive product scan of tf
#pragma unroll
              for (int sh = 1; sh < 32; sh <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, P, sh);
                if (lane >= sh) P = P * t;
              }
              float excl = __shfl_up_sync(0xffffffffu, P, 1);
              if (lane == 0) excl = 1.f;
              if (lane == 31) scan_s[qd] = P;
              bar128();
              float Wq = 1.f, Wall = 1.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float t = scan_s[q];
                if (q < qd) Wq = Wq * t;
                Wall = Wall * t;
              }
              const float Ti = (c_T * Wq) * excl;              // accum_prod (model_utils.py:110-113)
              const float w = al * Ti;
              const bool live = tile < num_tiles;              // (a cluster's surplus tile computes on clamped rows)
              if (args.ray_weights && live) args.ray_weights[row.m] = w;
              float C = w;                                     // inclusive cumsum of the weights
#pragma unroll
              for (int sh = 1; sh < 32; sh <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, C, sh);
                if (lane >= sh) C = C + t;
              }
              float red[6] = {w * o.x, w * o.y, w * o.z, w * row.z, w, row.last ? 0.f : w};
#pragma unroll
              for (int k = 0; k < 6; ++k)
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) red[k] += __shfl_xor_sync(0xffffffffu, red[k], sh);
              if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) scan_s[8 + qd * 8 + k] = red[k];
              }
              bar128();
              float pre = c_cw, tot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (q < qd) pre += scan_s[8 + q * 8 + 4];
#pragma unroll
                for (int k = 0; k < 6; ++k) tot[k] += scan_s[8 + q * 8 + k];
              }
              const float cw = pre + C;                        // cumsum over the ray up to this sample
              float prev = __shfl_up_sync(0xffffffffu, cw, 1);
              if (lane == 0) prev = pre;
              // first sample whose cumulative weight reaches 0.5 (opaque xor shifted, :231-238)
              float med = (cw >= 0.5f && !(prev >= 0.5f)) ? row.z : 0.f;
#pragma unroll
              for (int sh = 16; sh > 0; sh >>= 1) med += __shfl_xor_sync(0xffffffffu, med, sh);
              if (lane == 0) scan_s[4 + qd] = med;
              bar128();
              a_med += scan_s[4] + scan_s[5] + scan_s[6] + scan_s[7];
              a_r += tot[0]; a_g += tot[1]; a_b += tot[2]; a_d += tot[3]; a_w += tot[4]; a_wnl += tot[5];
              c_T = c_T * Wall; c_cw += tot[4];
              if ((tile % tpr) == tpr - 1 && r == 0 && live) {
                float rr = a_r, gg = a_g, bb = a_b;
                if (args.white_bg) { const float bg = 1.f - a_w; rr = rr + bg; gg = gg + bg; bb = bb + bg; }
                float* ro = args.ray_out + row.ray * 6;
                ro[0] = rr; ro[1] = gg; ro[2] = bb; ro[3] = a_d; ro[4] = a_med;
                ro[5] = args.sample_at_infinity ? a_wnl : a_w;
              }
            }
            if (next_stored) {
              row = nrow;                               // (arrive_all went out right after the accumulator was drained)
              next_stored = false;
            } else {
              if (has_next) {
                if (pf_step < 0) tile_prefetch(tile_of(ti + 1), pf);
                begin_tile(tile_of(ti + 1), pf);
              }
              arrive_all();
            }
            tr.ev(si, 5);
          }
        }
        if (encode_here) {
          // this step's last accumulator has been waited for: the input block has no reader left
          store_packed_x3(inh, inl, r, hs, nqh, nql);
          fence_proxy_async();
          next_stored = true;
        }
      }
    }
    tr.finish(args, 1 + hs);
    tc_fence_before();
  }
  if constexpr (kCl > 1) cluster_sync_all();  // nobody exits while a peer may still multicast into it / signal it
  else __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmCols);
}

// Cluster launch (-DNFB_X3_CLUSTER=2|4): the CTAs of a cluster share ONE weight stream - every rank
// fetches 1 / kCl of each slot and multicasts it into all rings, so the L2 -> SM weight traffic
// (2.8 MB per 128-row tile and SM, ~6 TB/s over the chip) drops by kCl; the CTAs are otherwise
// independent (own tiles, MMAs, TMEM) but a slot is refilled only once ALL issuers have released it.
#ifndef NFB_X3_CLUSTER
#define NFB_X3_CLUSTER 1
#endif
constexpr int kX3Cluster = NFB_X3_CLUSTER;
static_assert(kX3Cluster == 1 || kX3Cluster == 2 || kX3Cluster == 4, "cluster size");

inline int create_x3(nfb_handle*) {
  if (cudaFuncSetAttribute(field_x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kX3SmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(field_x3_kernel<kX3Cluster>, cudaFuncAttributeMaxDynamicSharedMemorySize, kX3SmemBytes) != cudaSuccess)
    return fail("cannot reserve %d bytes of shared memory for the fp16x3 kernel", kX3SmemBytes);
  return 0;
}

inline int run_field_x3(nfb_handle* h, int level, const FieldArgs& a, cudaStream_t s) {
  const long long tiles = (a.num_rows + kTileRows - 1) / kTileRows;
  if (tiles > 0x7fffffffLL) return fail("too many rows for one launch");
  const bool fuse = a.ray_out != nullptr && !a.warp_only;
  if (fuse && (a.samples_per_ray % kTileRows != 0 || a.num_rows % a.samples_per_ray != 0))
    return fail("fused composite needs samples_per_ray to be a multiple of %d", kTileRows);
  const long long groups = fuse ? a.num_rows / a.samples_per_ray : tiles;
  if (kX3Cluster == 1) h->x3_pair_ok = 0;
  if (groups >= (long long)h->sm_count && h->x3_pair_ok != 0) {
    // clusters sharing the weight stream (see above): worth it once every SM has work
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kX3Threads); cfg.dynamicSmemBytes = kX3SmemBytes; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kX3Cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (h->x3_pair_ok < 0) {
      // persistent kernel: no more clusters than can be co-resident
      cfg.gridDim = dim3((unsigned)(h->sm_count / kX3Cluster * kX3Cluster));
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, field_x3_kernel<kX3Cluster>, &cfg) != cudaSuccess || n < 1) {
        cudaGetLastError();
        h->x3_pair_ok = 0;
      } else {
        h->x3_pair_ok = n;
      }
    }
    if (h->x3_pair_ok > 0) {
      const int gridc = kX3Cluster * (int)std::min<long long>((groups + kX3Cluster - 1) / kX3Cluster, h->x3_pair_ok);
      cfg.gridDim = dim3((unsigned)gridc);
      cudaError_t le = cudaLaunchKernelEx(&cfg, field_x3_kernel<kX3Cluster>, h->tcprog[level], h->x3c[level], a,
                                          (const uint8_t*)h->d_wpack, (int)tiles, (const float*)h->d_aux);
      if (le != cudaSuccess) return fail("field_x3_kernel (cluster) launch failed: %s", cudaGetErrorString(le));
      h->launches++;
      return 0;
    }
  }
  const int grid = (int)std::min<long long>(groups, h->sm_count);
  field_x3_kernel<1><<<grid, kX3Threads, kX3SmemBytes, s>>>(h->tcprog[level], h->x3c[level], a, h->d_wpack, (int)tiles, (const float*)h->d_aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("field_x3_kernel launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  return 0;
}

}  // namespace tc3
}  // namespace nfb
