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
// Layout: ONE 128-row tile per CTA at a time (hi and lo images of the activations
// double the shared-memory footprint, so there is no room for the two sub-tiles of
// the bf16 kernel): activations hi [4][16 KB] | lo [4][16 KB] | input block hi | lo |
// weight ring 4 x 16 KB.  The 3x tensor work per layer hides the epilogue without a
// second tile: chunk 0's epilogue overlaps chunk 1's MMAs, chunk 1's epilogue
// overlaps the first K-blocks of the next layer (those blocks come from chunk 0).
// Warp roles (384 threads): warps 0-7 epilogue, TWO threads per row (warps w and
// w + 4 share a TMEM lane quarter and split a chunk's columns), warp 8 issues the
// MMAs, warp 9 streams the weights (cp.async.bulk, [W_hi | W_lo] per K-block and
// chunk), warps 10-11 complete the control warpgroup (setmaxnreg 40 / 232).
// One issuer "unit" = one K-block of one chunk = 12 MMAs (two ring stages).
#pragma once
#include <cuda_fp16.h>

#include "field_tc.cuh"

namespace nfb {
namespace tc3 {

using namespace nfb::tc;

constexpr int kX3Threads = 384;
constexpr int kX3EpiThreads = 256;
constexpr int kX3Stages = 4;
constexpr int kXlOff = 4 * kABlockBytes;                 // lo image of the activations
constexpr int kInhOff = 8 * kABlockBytes;                // == tc::kXBytes: the unit table's input-block offset
constexpr int kInlOff = 9 * kABlockBytes;
constexpr int kStageOff = 10 * kABlockBytes;
constexpr int kPartOff = kStageOff + kX3Stages * kStageBytes;   // alpha partial of the row's second thread
constexpr int kBarOff = kPartOff + 512;
constexpr int kX3SmemBytes = kBarOff + 128;
static_assert(kInhOff == kXBytes, "unit table offsets");
static_assert(kX3SmemBytes <= 232448, "shared memory");

struct X3Bars {
  uint64_t full[kX3Stages];
  uint64_t empty[kX3Stages];
  uint64_t acc_ready[2];
  uint64_t x_free;
  uint64_t x_ready[3];
  uint32_t tmem_slot;
};
static_assert(sizeof(X3Bars) <= 128, "barrier block");

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

// [x | w_k sin/cos features | extra | 0...] -> chunks [c_begin, c_end) of one
// 64-column K-block row (modules.py:213-228, 262-271; libm sinf as the fp32 path).
__device__ __forceinline__ void posenc_to_block_x3(uint8_t* bh, uint8_t* bl, int r, const float* x, int F,
                                                   const float* __restrict__ window,
                                                   const float* __restrict__ extra, int n_extra,
                                                   int c_begin, int c_end) {
  const int nf = 6 * F;
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c * 8 + j;
      float val = 0.f;
      if (k < 3) {
        val = x[k];
      } else if (k < 3 + nf) {
        const int f = k - 3;
        val = posenc_feature(x, f);
        if (window) val = __ldg(window + f / 6) * val;
      } else if (k < 3 + nf + n_extra) {
        val = __ldg(extra + (k - 3 - nf));
      }
      v[j] = val;
    }
    store_chunk_x3(bh, bl, r, c, v);
  }
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

// One 32-column piece of a hidden layer's epilogue: + bias, activation, (alpha head
// dot product in fp32), split into fp16 hi / lo pairs.
// `inv_s` undoes the power-of-two weight scale of the layer (x3_weight_scale): acc * inv_s is
// exact, so fma(acc, inv_s, bias) rounds exactly like the unscaled acc + bias.
__device__ __forceinline__ void x3_piece(const float* v, const float4* __restrict__ bq4, float inv_s, bool relu,
                                         bool adot, const float4* __restrict__ aw4, float& alpha,
                                         uint32_t* hi16, uint32_t* lo16) {
  const uint64_t is2 = pack_f32x2(inv_s, inv_s);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 bq = bq4[j >> 2];                  // constant bank, warp-uniform address
    uint64_t r0, r1;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r0) : "l"(pack_f32x2(v[j], v[j + 1])), "l"(is2), "l"(pack_f32x2(bq.x, bq.y)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r1) : "l"(pack_f32x2(v[j + 2], v[j + 3])), "l"(is2), "l"(pack_f32x2(bq.z, bq.w)));
    float t0, t1, t2, t3;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(r0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t2), "=f"(t3) : "l"(r1));
    if (relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f); }
    if (adot) {
      const float4 w = aw4[j >> 2];
      alpha = fmaf(t0, w.x, alpha); alpha = fmaf(t1, w.y, alpha);
      alpha = fmaf(t2, w.z, alpha); alpha = fmaf(t3, w.w, alpha);
    }
    split_pair(t0, t1, hi16[j >> 1], lo16[j >> 1]);
    split_pair(t2, t3, hi16[(j >> 1) + 1], lo16[(j >> 1) + 1]);
  }
}

// Second and third MMA chains of a unit plus everything that follows them:
//   4 x (x_lo W_hi), commit "stage s free", 4 x (x_hi W_lo) with the look-ahead
//   probes of the next unit's barriers in between, commit "stage s+1 free" and the
//   optional x_free / accumulator commits.  Returns the probe bits:
//   1 = both weight stages of the next unit have landed, 2/4/8 = x_ready[0/1/2].
__device__ __forceinline__ uint32_t issue_x3_tail(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                                  uint64_t b_lo, uint32_t idesc, uint32_t bar_e0,
                                                  uint32_t bar_e1, uint32_t bar_xfree, uint32_t bar_acc,
                                                  uint32_t probe_w0, uint32_t probe_w1, uint32_t par_w,
                                                  uint32_t probe_x0, uint32_t probe_x1, uint32_t probe_x2,
                                                  uint32_t par_x) {
  uint32_t out;
  asm volatile(
      "{\n\t"
      ".reg .pred pt, pw0, pw1, px0, px1, px2, pd0, pd1, pd2, pcx, pca;\n\t"
      ".reg .b64 l1, l2, l3, h1, h2, h3, bh1, bh2, bh3, bl1, bl2, bl3;\n\t"
      ".reg .b32 t0, t1, t2;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "setp.ne.b32 pd0, %14, 0;\n\t"
      "setp.ne.b32 pd1, %15, 0;\n\t"
      "setp.ne.b32 pd2, %16, 0;\n\t"
      "setp.ne.b32 pcx, %9, 0;\n\t"
      "setp.ne.b32 pca, %10, 0;\n\t"
      "setp.eq.b32 px0, 1, 0;\n\t"
      "setp.eq.b32 px1, 1, 0;\n\t"
      "setp.eq.b32 px2, 1, 0;\n\t"
      "add.u64 l1, %3, 2;\n\t add.u64 l2, %3, 4;\n\t add.u64 l3, %3, 6;\n\t"
      "add.u64 h1, %2, 2;\n\t add.u64 h2, %2, 4;\n\t add.u64 h3, %2, 6;\n\t"
      "add.u64 bh1, %4, 2;\n\t add.u64 bh2, %4, 4;\n\t add.u64 bh3, %4, 6;\n\t"
      "add.u64 bl1, %5, 2;\n\t add.u64 bl2, %5, 4;\n\t add.u64 bl3, %5, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %4, %6, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], l1, bh1, %6, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], l2, bh2, %6, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], l3, bh3, %6, pt;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %5, %6, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], h1, bl1, %6, pt;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 pw0, [%11], %13;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 pw1, [%12], %13;\n\t"
      "@pd0 mbarrier.test_wait.parity.shared::cta.b64 px0, [%14], %17;\n\t"
      "@pd1 mbarrier.test_wait.parity.shared::cta.b64 px1, [%15], %17;\n\t"
      "@pd2 mbarrier.test_wait.parity.shared::cta.b64 px2, [%16], %17;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], h2, bl2, %6, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], h3, bl3, %6, pt;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
      "@pcx tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
      "@pca tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
      "and.pred pw0, pw0, pw1;\n\t"
      "selp.u32 %0, 1, 0, pw0;\n\t"
      "selp.u32 t0, 2, 0, px0;\n\t"
      "selp.u32 t1, 4, 0, px1;\n\t"
      "selp.u32 t2, 8, 0, px2;\n\t"
      "or.b32 %0, %0, t0;\n\t"
      "or.b32 %0, %0, t1;\n\t"
      "or.b32 %0, %0, t2;\n\t"
      "}"
      : "=r"(out)
      : "r"(d), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(bar_e0), "r"(bar_e1),
        "r"(bar_xfree), "r"(bar_acc), "r"(probe_w0), "r"(probe_w1), "r"(par_w), "r"(probe_x0),
        "r"(probe_x1), "r"(probe_x2), "r"(par_x)
      : "memory");
  return out;
}

// Row state owned by the two epilogue threads of a row for the lifetime of a tile.
struct X3Row {
  float x[3];        // current (possibly warped) sample point
  long long m;       // global row (clamped to a valid row)
  long long ray;
  bool valid;
  float alpha;       // this thread's part of the alpha-head dot product
};

__global__ void __launch_bounds__(kX3Threads, 1)
field_x3_kernel(const __grid_constant__ TcProgram prog, const __grid_constant__ X3Consts cst,
                const FieldArgs args, const uint8_t* __restrict__ wpack, int num_tiles) {
  constexpr int kEpiWarps = 8, kMmaWarp = 8, kProdWarp = 9;
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* base = raw;
  if ((smem_u32(base) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("nfb: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* xh = base;                          // [4][16 KB] activations, hi
  uint8_t* inh = base + kInhOff;               // input block, hi
  uint8_t* inl = base + kInlOff;
  uint8_t* stages = base + kStageOff;
  float* alpha_part = reinterpret_cast<float*>(base + kPartOff);
  X3Bars* bars = reinterpret_cast<X3Bars*>(base + kBarOff);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == kMmaWarp * 32) {
    for (int i = 0; i < kX3Stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->acc_ready[0], 1); mbar_init(&bars->acc_ready[1], 1);
    mbar_init(&bars->x_free, 1);
    for (int k = 0; k < 3; ++k) mbar_init(&bars->x_ready[k], kX3EpiThreads);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&bars->tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
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

  if (warp == kProdWarp) {
    // ===================== weight producer =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
    uint32_t it = 0, dead = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int si = first_step; si <= last_step; ++si) {
        const TcStep& st = prog.steps[si];
        const uint32_t bytes = (uint32_t)st.chunk_n * kRowBytes;
        const uint8_t* src = wpack + st.w_off;
        const int n = st.n_chunks * st.nkb * 2;           // [chunk][kb][W_hi, W_lo]
        for (int u = 0; u < n; ++u, ++it) {
          const int sg = it % kX3Stages;
          const uint32_t ph = (it / kX3Stages) & 1;
          mbar_wait(&bars->empty[sg], ph ^ 1, dead);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bars->full[sg], bytes);
            bulk_g2s(stages + sg * kStageBytes, src + (size_t)u * bytes, bytes, &bars->full[sg]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
    if (elect_one()) {
      Tracer tr(args, 0);
      const uint64_t desc_hi = make_smem_desc(0) & 0xFFFFFFFF00000000ull;
      const uint32_t lo_base = ((smem_u32(xh) & 0x3FFFFu) >> 4) | (1u << 16);
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
      uint64_t ad0 = desc_hi | (uint64_t)(lo_base + c0.x);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int u = u_begin; u < u_end; ++u) {
          const uint32_t flags = c1.x, need = c1.y;
          if ((ready & need) != need) {                    // slow path: something is not there yet
            if ((need & 2) && !(ready & 2)) mbar_wait_issuer(&bars->x_ready[0], xr & 1, dead);
            if ((need & 4) && !(ready & 4)) mbar_wait_issuer(&bars->x_ready[1], xr & 1, dead);
            if ((need & 8) && !(ready & 8)) mbar_wait_issuer(&bars->x_ready[2], xr & 1, dead);
            if (!(ready & 1)) {
              mbar_wait_issuer(&bars->full[sg], wph, dead);
              mbar_wait_issuer(&bars->full[sg + 1], wph, dead);
            }
          }
          issue_half0(d0, ad0, bd, c0.w, flags & kUAccum);               // x_hi W_hi
          // ---- bookkeeping while those MMAs execute ----
          if (++un >= u_end) un -= n_u;
          const uint4 f0 = utab[2 * un], f1 = utab[2 * un + 1];          // table entry two units ahead
          const uint32_t nsg = (sg + 2) & (kX3Stages - 1);
          const uint32_t nwph = wph ^ (nsg == 0 ? 1u : 0u);
          const uint32_t nxr = xr + ((flags & kUStepEnd) ? 1u : 0u);
          const uint64_t a_lo = desc_hi | (uint64_t)(lo_base + c0.y);
          const uint32_t px0 = (c1.z & 2) ? b_x0 : 0u, px1 = (c1.z & 4) ? b_x1 : 0u;
          const uint32_t px2 = (c1.z & 8) ? b_x2 : 0u;
          const uint32_t d_cur = d0, idesc = c0.w, bar_e = b_empty + sg * 8;
          const uint64_t bd_cur = bd, a_hi = ad0;
          d0 = tmem_base + n0.z;
          bd = desc_hi | (uint64_t)(st_lo + nsg * (kStageBytes >> 4));
          ad0 = desc_hi | (uint64_t)(lo_base + n0.x);
          ready = issue_x3_tail(d_cur, a_hi, a_lo, bd_cur, bd_cur + (uint64_t)(kStageBytes >> 4), idesc,
                                bar_e, bar_e + 8, (flags & kUCommitXFree) ? b_xfree : 0u,
                                (flags & kUCommitAcc0) ? b_acc0 : ((flags & kUCommitAcc1) ? b_acc1 : 0u),
                                b_full + nsg * 8, b_full + nsg * 8 + 8, nwph, px0, px1, px2, nxr & 1);
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
    const uint32_t xs_a0 = smem_u32(xh) + r * kRowBytes + ((r & 7) << 4);   // see sts_piece()
    const int cb = hs * 4, ce = cb + 4;                // input-block chunks this thread writes
    Tracer tr(args, (lane == 0 && qd == 0) ? 1 + hs : -1);
    uint32_t n_acc0 = 0, n_acc1 = 0, n_free = 0, dead = 0;
    X3Row row;
    const int S = args.samples_per_ray;
    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kX3EpiThreads) : "memory"); };
    auto arrive_all = [&]() {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars->x_ready[0]);
      mbar_arrive(&bars->x_ready[1]);
      mbar_arrive(&bars->x_ready[2]);
    };
    // Sample point of this thread's row in tile `tile`, and the first input block
    // (model_utils.py:72-73; warping.py:325-326 / models.py:270).
    auto begin_tile = [&](int tile) {
      long long m = (long long)tile * kTileRows + r;
      row.valid = m < args.num_rows;
      if (!row.valid) m = args.num_rows - 1;
      row.m = m;
      row.ray = m / S;
      const float z = args.z_vals ? __ldg(args.z_vals + m) : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        row.x[c] = __ldg(args.origins + row.ray * 3 + c) + z * __ldg(args.directions + row.ray * 3 + c);
      const float* cond = args.cond + row.ray * prog.cond_stride;
      if (do_warp) {
        posenc_to_block_x3(inh, inl, r, row.x, prog.Fw, args.window, cond, prog.G, cb, ce);
      } else {
        if (args.warped && row.valid && hs == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) args.warped[m * 3 + c] = row.x[c];
        }
        posenc_to_block_x3(inh, inl, r, row.x, prog.Fp, nullptr, nullptr, 0, cb, ce);
      }
      row.alpha = hs == 0 ? cst.alpha_b : 0.f;
    };

    int tile = blockIdx.x;
    if (tile < num_tiles) {
      begin_tile(tile);
      arrive_all();
    }
    for (; tile < num_tiles; tile += gridDim.x) {
      for (int si = first_step; si <= last_step; ++si) {
        const TcStep& st = prog.steps[si];
        const float4* bias4 = cst.b4 + si * 64;        // this step's 256 biases
        const float inv_s = cst.inv_scale[si];          // undoes the step's power-of-two weight scale
        if (st.epi == kEpiHidden) {
          const int cols = st.chunk_n >> 1;            // columns of a chunk handled by this thread: 64 or 32
          const bool wide = cols == 64;
          const bool relu = st.relu != 0, adot = st.alpha_dot != 0;
          // ---- chunk 0: results wait in registers until the MMAs of chunk 1 no
          //      longer read the blocks they overwrite ----
          uint32_t ph[32], pl[32];
          const int col0 = hs * cols;
          mbar_wait(&bars->acc_ready[0], n_acc0++ & 1, dead);
          tc_fence_after();
          tr.ev(si, 0);
          if (wide) {
            float va[32], vb[32];
            tmem_ld32(t_lane + col0, va);
            tmem_ld32(t_lane + col0 + 32, vb);
            tmem_ld_wait();
            x3_piece(va, bias4 + (col0 >> 2), inv_s, relu, adot, cst.alpha4 + (col0 >> 2), row.alpha, ph, pl);
            x3_piece(vb, bias4 + (col0 >> 2) + 8, inv_s, relu, adot, cst.alpha4 + (col0 >> 2) + 8, row.alpha, ph + 16, pl + 16);
          } else {
            float va[32];
            tmem_ld32(t_lane + col0, va);
            tmem_ld_wait();
            x3_piece(va, bias4 + (col0 >> 2), inv_s, relu, adot, cst.alpha4 + (col0 >> 2), row.alpha, ph, pl);
          }
          tr.ev(si, 1);
          mbar_wait(&bars->x_free, n_free++ & 1, dead);
          tr.ev(si, 2);
          sts_piece(xs_a0, col0, ph);
          sts_piece(xs_a0 + kXlOff, col0, pl);
          if (wide) {
            sts_piece(xs_a0, col0 + 32, ph + 16);
            sts_piece(xs_a0 + kXlOff, col0 + 32, pl + 16);
          }
          // The input block is the first K-block of the next step (read right after
          // x_ready[0]); every earlier reader of it (the skip layer) is complete.
          if (st.write_cond)
            cond_to_block_x3(inh, inl, r, args.cond + row.ray * prog.cond_stride + prog.G, prog.rc, cb, ce);
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&bars->x_ready[0]);
          tr.ev(si, 3);
          // ---- chunk 1: every MMA of the layer is complete, store directly ----
          const int col1 = st.chunk_n + col0;
          mbar_wait(&bars->acc_ready[1], n_acc1++ & 1, dead);
          tc_fence_after();
          tr.ev(si, 4);
          if (wide) {
            float va[32], vb[32];
            tmem_ld32(t_lane + col1, va);
            tmem_ld32(t_lane + col1 + 32, vb);
            tmem_ld_wait();
            x3_piece(va, bias4 + (col1 >> 2), inv_s, relu, adot, cst.alpha4 + (col1 >> 2), row.alpha, ph, pl);
            sts_piece(xs_a0, col1, ph);
            sts_piece(xs_a0 + kXlOff, col1, pl);
            x3_piece(vb, bias4 + (col1 >> 2) + 8, inv_s, relu, adot, cst.alpha4 + (col1 >> 2) + 8, row.alpha, ph, pl);
            sts_piece(xs_a0, col1 + 32, ph);
            sts_piece(xs_a0 + kXlOff, col1 + 32, pl);
          } else {
            float va[32];
            tmem_ld32(t_lane + col1, va);
            tmem_ld_wait();
            x3_piece(va, bias4 + (col1 >> 2), inv_s, relu, adot, cst.alpha4 + (col1 >> 2), row.alpha, ph, pl);
            sts_piece(xs_a0, col1, ph);
            sts_piece(xs_a0 + kXlOff, col1, pl);
          }
          if (adot && hs == 1) alpha_part[r] = row.alpha;   // read by the row's first thread at the rgb step
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&bars->x_ready[1]);
          mbar_arrive(&bars->x_ready[2]);
          tr.ev(si, 5);
        } else {
          // ---- heads: N = 16 accumulator columns, one chunk (both threads of a row
          //      do the scalar work; they split the input-block chunks) ----
          float v[16];
          mbar_wait(&bars->acc_ready[0], n_acc0++ & 1, dead);
          tc_fence_after();
          tr.ev(si, 0);
          tmem_ld16(t_lane, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 12; j += 4) {
            const float4 bq = bias4[j >> 2];
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
              const int nxt = tile + gridDim.x;
              if (nxt < num_tiles) begin_tile(nxt);
            } else {
              posenc_to_block_x3(inh, inl, r, row.x, prog.Fp, nullptr, nullptr, 0, cb, ce);
            }
            arrive_all();
          } else {
            epi_sync();                                 // the partner's alpha partial is visible
            if (row.valid && args.samples && hs == 0) {
              float4 o;
              o.x = sigmoidf(v[0]); o.y = sigmoidf(v[1]); o.z = sigmoidf(v[2]);
              o.w = apply_act(row.alpha + alpha_part[r], prog.sigma_act);
              reinterpret_cast<float4*>(args.samples)[row.m] = o;
            }
            // (the next write of alpha_part[r] is ordered behind this read by the barrier
            // chain of the next tile's layers: x_ready <- this thread, acc_ready -> writer)
            const int nxt = tile + gridDim.x;
            if (nxt < num_tiles) begin_tile(nxt);
            arrive_all();
            tr.ev(si, 5);
          }
        }
      }
    }
    tr.finish(args, 1 + hs);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 256);
}

inline int create_x3(nfb_handle*) {
  if (cudaFuncSetAttribute(field_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kX3SmemBytes) != cudaSuccess)
    return fail("cannot reserve %d bytes of shared memory for the fp16x3 kernel", kX3SmemBytes);
  return 0;
}

inline int run_field_x3(nfb_handle* h, int level, const FieldArgs& a, cudaStream_t s) {
  const long long tiles = (a.num_rows + kTileRows - 1) / kTileRows;
  if (tiles > 0x7fffffffLL) return fail("too many rows for one launch");
  const int grid = (int)std::min<long long>(tiles, h->sm_count);
  field_x3_kernel<<<grid, kX3Threads, kX3SmemBytes, s>>>(h->tcprog[level], h->x3c[level], a, h->d_wpack, (int)tiles);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("field_x3_kernel launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  return 0;
}

}  // namespace tc3
}  // namespace nfb
