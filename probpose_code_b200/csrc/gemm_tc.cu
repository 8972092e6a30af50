// tcgen05 tensor-core GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] . W[N,K]^T), fp32 accumulate.
//
// Persistent, warp-specialised: one CTA per SM walks the 128 x BN output tiles (n fastest, so the
// CTAs running together share their A rows in L2) and keeps three pipelines busy at once:
//   warp 0      TMA producer  - cp.async.bulk.tensor (128-byte swizzle) of A / W k-blocks into a
//                               STAGES-deep shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer    - one thread issues tcgen05.mma (M=128, N=BN, K=16) per 32-byte
//                               k-slice into one of TWO TMEM accumulator stages; tcgen05.commit
//                               frees the smem slot / publishes the accumulator
//   warps 2..9  epilogue      - drain the other accumulator stage while the next tile's MMAs run:
//                               tcgen05.ld (32 lanes x 32 columns) -> registers -> swizzled smem
//                               transpose -> scale/shift/activation/residual on float4 row segments
//                               -> coalesced global stores (fp32 rows, next operand, or planes)
//
// Precision modes (pp_precision):
//   FP16 / BF16  one MMA per k-slice.
//   FP16X3       operands are [hi | lo] fp16 pairs of the 64x-scaled values (common.cuh); three MMAs
//                per k-slice into ONE accumulator: acc += Ahi.Whi + Ahi.Wlo + Alo.Whi.
//                fp16 products are exact in the fp32 accumulator, so this recovers ~2^-22
//                relative operand precision (fp32-grade) at 1/3 of the fp16 tensor rate; the
//                2^12 operand scale is removed exactly in the epilogue.
//
// Tile width: L2 -> shared-memory traffic per MAC goes as (1/BM + 1/BN), and with hi + lo planes the
// 128 x 128 tile is bound by the ~6 KB/clk L2 fabric at ~47 % tensor activity (measured); the single
// accumulator lets FP16X3 run 128 x 192 / 128 x 256 tiles with both TMEM stages (2 x 256 columns).
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace pp {

constexpr int kBM = 128;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;  // TMA warp + MMA warp + epilogue warps
constexpr int kStagingBytes = kEpiWarps * 4096;    // one 32 x 32 fp32 transpose tile per epilogue warp
constexpr int kSmemMax = 227 * 1024;
constexpr int kMaxTaps = 9;

// Implicit-GEMM A operand ("taps"): K is split into `taps` groups of `tap_k` columns; group t is
// read from the SAME source operand (row width tap_k) at row offset shift[t].  The source rows
// enumerate a zero-padded NHWC map, so a constant row shift is a spatial (dy, dx) shift and
// conv / deconv-phase GEMMs need no gathered copy of their input.  taps == 1: plain operand.
struct TapParams {
  int taps;
  int tap_k;
  int shift[kMaxTaps];
};

// Grouped launch: up to kMaxGroups GEMMs that share A, the output buffer, the shapes and the epilogue - the four
// sub-pixel phases of a ConvTranspose2d(k4, s2, p1) - run as ONE persistent launch.  Group 0 is described by the plain
// arguments; groups 1.. bring their own W operand, tap shifts and phase.  Tiles are ordered (m, group, n), so the phases of
// a row block run side by side (one pass over the A rows in L2), the launch has one fill / drain instead of four and the
// last round of the tile walk is shared (deconv-2 at batch 64: 4 x 826 tiles on 74 CTA pairs, 45 rounds instead of 4 x 12).
constexpr int kMaxGroups = 4;
struct GroupParams {
  CUtensorMap tm_w[kMaxGroups - 1];
  int groups;  // 1 = plain launch
  int shift[kMaxGroups - 1][kMaxTaps];
  int up_py[kMaxGroups - 1], up_px[kMaxGroups - 1];
};

// ACCS (FP16X3 only): 1 = all three products accumulate into one TMEM accumulator (wide tiles keep
// both TMEM stages); 2 = the cross terms hi.lo + lo.hi go to a second accumulator, so the big
// accumulator is rounded once per k-slice instead of three times (used for long K, see pick_config).
// PAIR: two CTAs of a cluster (one TPC) form one 256 x BN tile with tcgen05 cta_group::2: each CTA
// stages its own 128 A rows and HALF of the W rows, the leader issues M = 256 MMAs that read both
// shared memories, and each CTA's TMEM receives its 128 accumulator rows.  Per CTA this halves the W
// bytes pulled from L2 and read from shared memory - the two limits a 128-row tile runs into first
// (measured: 47 % tensor activity for 128 x 128 FP16X3 tiles, L2 -> SM traffic at 10 TB/s).
template <int BN, int SPLIT, int ACCS = 1, bool PAIR = false>
struct GemmCfg {
  // k-block: 64 x 16-bit = one 128-byte swizzle row; wide FP16X3 tiles use 32 (64-byte swizzle rows)
  // so that four stages of A hi/lo + W hi/lo still fit beside the epilogue staging tiles
  static constexpr int BK = (SPLIT == 3 && BN > 128 && !PAIR) ? 32 : 64;
  static constexpr int BN_CTA = PAIR ? BN / 2 : BN;  // W rows this CTA stages
  static constexpr int NOPS = (SPLIT == 3) ? 2 : 1;  // operand planes (hi, lo)
  static constexpr int A_BYTES = kBM * BK * 2;
  static constexpr int B_BYTES = BN_CTA * BK * 2;
  static constexpr int STAGE_BYTES = NOPS * (A_BYTES + B_BYTES);
  static constexpr int ACC_COLS = ACCS * BN;
  static constexpr int ACC_STAGES = (2 * ACC_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_NEED = ACC_STAGES * ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int MISC_BYTES = 256;  // barriers + tmem pointer
  static constexpr int RING_BUDGET = kSmemMax - 1024 - kStagingBytes - MISC_BYTES;
  static constexpr int STAGES_RAW = RING_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  // ring | staging | full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2] | tmem ptr | 1024 B alignment slack
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kStagingBytes + MISC_BYTES + 1024;
  static_assert(STAGES >= 2, "tile too large for a 2-stage ring");
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert((2 * STAGES + 4) * 8 + 16 <= MISC_BYTES, "barrier block too small");
  static_assert(BN % 32 == 0, "epilogue works on 32-column chunks");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must keep the swizzle alignment");
  static_assert(!PAIR || BN % 32 == 0, "pair tiles split W in two halves of whole 16-row groups");
};

// erf by Abramowitz & Stegun 7.1.26 (|error| < 6e-7 in fp32, below the rounding noise of an fp32
// GELU): branch-free, two MUFU ops, so the fc1 epilogue stays under the tile's MMA time and the
// kernel stays inside the instruction cache.  (The CUDA-core verification GEMM keeps erff.)
// Returns OUT_SCALE * gelu(v); the operand scale of the next GEMM rides along for free:
// s * 0.5 v (1 + erf) = h + |h| E with h = 0.5 s v, E = erf(|x|) (x and h share their sign).
template <int OUT_SCALE>
__device__ __forceinline__ float gelu_fast(float v) {
  const float a = fabsf(v) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((a * -1.4426950408889634f) * a));
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float h = v * (0.5f * OUT_SCALE);
  return fmaf(fabsf(h), erf_abs, h);
}

// The same GELU on two values at once with the packed fp32 instructions of sm_100 (FFMA2: two IEEE fp32 FMAs per
// issue slot).  The fc1 epilogue is bound by issue slots (about 21 per element against 9.2 K MMA cycles per
// 128 x 256 tile); the polynomial, the exponent argument and the final blend take 12 packed instructions per PAIR
// instead of 12 per element.  Same operation order as gelu_fast, lane for lane (the polynomial carries the sign
// of -p in its coefficients; |h| erf = h copysign(erf, v)): bit-identical results.
// (x0, x1) <- OUT_SCALE * gelu(x * sc + sh) for two neighbouring columns
template <int OUT_SCALE>
__device__ __forceinline__ void affine_gelu2(float& x0, float& x1, uint64_t sc, uint64_t sh) {
  const uint64_t v = fma2(pk2(x0, x1), sc, sh);
  float v0, v1;
  upk2(v, v0, v1);
  const uint64_t a = mul2(pk2(fabsf(v0), fabsf(v1)), bc2(0.70710678118654752440f));
  const uint64_t d = fma2(bc2(0.3275911f), a, bc2(1.0f));
  float d0, d1, t0, t1, g0, g1, e0, e1;
  upk2(d, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t = pk2(t0, t1);
  uint64_t np = fma2(bc2(-1.061405429f), t, bc2(1.453152027f));  // -p: negated coefficients
  np = fma2(np, t, bc2(-1.421413741f));
  np = fma2(np, t, bc2(0.284496736f));
  np = fma2(np, t, bc2(-0.254829592f));
  np = mul2(np, t);
  upk2(mul2(mul2(a, bc2(-1.4426950408889634f)), a), g0, g1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const uint64_t erf_abs = fma2(np, pk2(e0, e1), bc2(1.0f));
  float r0, r1;
  upk2(erf_abs, r0, r1);
  const uint64_t h = mul2(v, bc2(0.5f * OUT_SCALE));
  upk2(fma2(h, pk2(copysignf(r0, v0), copysignf(r1, v1)), h), x0, x1);
}

__device__ __forceinline__ void st_shared_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_shared_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// Four consecutive columns of one operand row (col % 4 == 0), see common.cuh "GEMM operand formats";
// rowk = row * k (element offset of the row in the LOGICAL (rows, k) matrix).  FP16X3: v is already
// in operand units (64x the value, folded into the epilogue's affine step).
template <int PREC>
__device__ __forceinline__ void store_operand4_row(void* base, int64_t rowk, int col, int k, float4 v) {
  if constexpr (PREC == PP_PREC_FP16X3) {
    uint2 hi, lo;
#ifndef PP_NO_GEMM_GUARD
    note_overflow4(v.x, v.y, v.z, v.w);  // the conversion below clamps: never silently
#endif
    hi.x = pack_half2_sat(v.x, v.y); hi.y = pack_half2_sat(v.z, v.w);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
    float r0, r1, r2, r3;  // residues v - hi: exact, one packed subtraction per pair
    upk2(sub2(pk2(v.x, v.y), pk2(f0.x, f0.y)), r0, r1);
    upk2(sub2(pk2(v.z, v.w), pk2(f1.x, f1.y)), r2, r3);
    const __half2 l0 = __floats2half2_rn(r0, r1);
    const __half2 l1 = __floats2half2_rn(r2, r3);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
    __half* p = reinterpret_cast<__half*>(base) + 2 * rowk + col;
    *reinterpret_cast<uint2*>(p) = hi;
    *reinterpret_cast<uint2*>(p + k) = lo;
  } else if constexpr (PREC == PP_PREC_BF16) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + rowk + col) = o;
  } else {
    uint2 o;
    note_overflow4(v.x, v.y, v.z, v.w);
    o.x = pack_half2_sat(v.x, v.y); o.y = pack_half2_sat(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(base) + rowk + col) = o;
  }
}

// num_m_tiles counts 128-row tiles (PAIR: 256-row pair tiles); gridDim.x CTAs (PAIR: an even number,
// launched as clusters of 2) walk them persistently.
template <int BN, int SPLIT, bool BF16, int OUT, int ACCS, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const int K,
               const int num_m_tiles, const int num_n_tiles, const TapParams tp, const EpiParams e,
               const __grid_constant__ GroupParams gp) {
  using Cfg = GemmCfg<BN, SPLIT, ACCS, PAIR>;
  static_assert(ACCS == 1 || SPLIT == 3, "a second accumulator only exists in the split mode");
  constexpr int PREC = SPLIT == 3 ? PP_PREC_FP16X3 : (BF16 ? PP_PREC_BF16 : PP_PREC_FP16);
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float4* staging = reinterpret_cast<float4*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + kStagingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  auto stage_a = [&](int s, int part) { return smem + s * Cfg::STAGE_BYTES + part * Cfg::A_BYTES; };
  auto stage_b = [&](int s, int part) { return smem + s * Cfg::STAGE_BYTES + Cfg::NOPS * Cfg::A_BYTES + part * Cfg::B_BYTES; };

  constexpr int BK = Cfg::BK;
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = K / BK;
  const int G = gp.groups;
  const int gn_tiles = G * num_n_tiles;  // tiles per row block: (group, n)
  const int num_tiles = num_m_tiles * gn_tiles;
  // work-unit walk: a CTA (or a CTA pair) starts at its index and strides by the number of units
  const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;  // 0 = leader of the pair
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = PAIR ? 2 * kBM : kBM;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_w);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full_bar[s], 1);
      ptx::mbar_init(&tmem_empty_bar[s], PAIR ? 2 * kEpiWarps : kEpiWarps);  // the leader collects both CTAs' epilogues
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_2cta(tmem_ptr, Cfg::TMEM_COLS);
      ptx::tmem_relinquish_2cta();
    } else {
      ptx::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tcgen05_fence_before();
  if constexpr (PAIR) ptx::cluster_sync(); else __syncthreads();  // barriers + TMEM visible (pair: in both CTAs)
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // everything above overlapped the previous kernel's tail; operands, residuals and outputs are only
  // touched after it has completed
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int kb_per_tap = tp.tap_k / BK;
      const int a_lo = tp.tap_k;  // column of the lo plane inside an FP16X3 A row
      uint32_t it = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
        const int m0 = (tile / gn_tiles) * TILE_M + (int)cta_rank * kBM;
        const int gi = (tile % gn_tiles) / num_n_tiles;  // group: its own W operand and tap shifts
        const int n0 = (tile % num_n_tiles) * BN + (int)cta_rank * Cfg::BN_CTA;  // pair: this CTA's half of the W tile
        const CUtensorMap* tmw = gi == 0 ? &tm_w : &gp.tm_w[gi - 1];
        const int* shifts = gi == 0 ? tp.shift : gp.shift[gi - 1];
        int tap = 0, kc = 0;  // current tap and its k-block
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          const int arow = m0 + shifts[tap];
          if constexpr (PAIR) {
            // both CTAs' bytes are credited to the leader's barrier; only the leader arms it
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
            ptx::tma_load_2d_2cta(stage_a(s, 0), &tm_a, &full_bar[s], kc * BK, arow);
            ptx::tma_load_2d_2cta(stage_b(s, 0), tmw, &full_bar[s], kb * BK, n0);
            if constexpr (SPLIT == 3) {
              ptx::tma_load_2d_2cta(stage_a(s, 1), &tm_a, &full_bar[s], a_lo + kc * BK, arow);
              ptx::tma_load_2d_2cta(stage_b(s, 1), tmw, &full_bar[s], K + kb * BK, n0);
            }
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
            ptx::tma_load_2d(stage_a(s, 0), &tm_a, &full_bar[s], kc * BK, arow);
            ptx::tma_load_2d(stage_b(s, 0), tmw, &full_bar[s], kb * BK, n0);
            if constexpr (SPLIT == 3) {
              ptx::tma_load_2d(stage_a(s, 1), &tm_a, &full_bar[s], a_lo + kc * BK, arow);
              ptx::tma_load_2d(stage_b(s, 1), tmw, &full_bar[s], K + kb * BK, n0);
            }
          }
          if (++kc == kb_per_tap) { kc = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ===== MMA issuer: the whole warp walks the pipeline (warp-uniform control flow keeps the
    // descriptors in uniform registers); one elected lane issues the MMAs and commits.
    // Pair mode: only the leader CTA issues (M = 256 across both CTAs) =====
    constexpr uint32_t idesc = ptx::make_idesc_f16(BF16, TILE_M, BN);
    constexpr uint32_t A_STEP = Cfg::A_BYTES >> 4, B_STEP = Cfg::B_BYTES >> 4;  // plane strides, 16-byte units
    const uint32_t desc_lo0 = ptx::kmajor_desc_lo(ptx::smem_u32(smem));
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t accumulate) {
      if constexpr (PAIR) ptx::umma_f16_2cta(d, da, db, idesc, accumulate); else ptx::umma_f16(d, da, db, idesc, accumulate);
    };
    auto commit = [&](uint64_t* bar) {  // pair: the arrival lands in both CTAs
      if constexpr (PAIR) ptx::umma_commit_2cta(bar); else ptx::umma_commit(bar);
    };
    uint32_t it = 0, lt = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++lt) {
      const int as = lt % ACC_STAGES;
      const uint32_t aph = (lt / ACC_STAGES) & 1;
      ptx::mbar_wait(&tmem_empty_bar[as], aph ^ 1);  // epilogue has drained this accumulator stage
      ptx::tcgen05_fence_after();
      const uint32_t acc0 = tmem_base + as * Cfg::ACC_COLS;
      const uint32_t acc1 = acc0 + (ACCS == 2 ? BN : 0);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tcgen05_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_hi = desc_lo0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
          const uint32_t b_hi = a_hi + Cfg::NOPS * A_STEP;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
            const uint64_t da = ptx::kmajor_desc<BK * 2>(a_hi + 2 * kk);
            const uint64_t db = ptx::kmajor_desc<BK * 2>(b_hi + 2 * kk);
            mma(acc0, da, db, acc);
            if constexpr (SPLIT == 3) {
              const uint64_t da_lo = ptx::kmajor_desc<BK * 2>(a_hi + A_STEP + 2 * kk);
              const uint64_t db_lo = ptx::kmajor_desc<BK * 2>(b_hi + B_STEP + 2 * kk);
              mma(acc1, da, db_lo, ACCS == 2 ? acc : 1u);
              mma(acc1, da_lo, db, 1u);
            }
          }
          commit(&empty_bar[s]);                                // slot reusable once these MMAs have read it
          if (kb == num_kb - 1) commit(&tmem_full_bar[as]);     // accumulator stage complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2) {
    // ===== epilogue (8 warps): two warps share a TMEM lane quarter and split its column chunks =====
    const int ew = warp - 2;
    const int q = warp & 3;     // TMEM lane quarter this warp may access
    const int half = ew >> 2;   // which of the quarter's two warps
    const uint32_t stg = ptx::smem_u32(staging) + ew * 4096;  // this warp's 32 x 32 fp32 transpose tile
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int c4 = lane & 7, rsub = lane >> 3;
    const bool vec_ok = (e.ldd & 3) == 0;
    constexpr float acc_scale = SPLIT == 3 ? kAccScaleInv : 1.0f;  // FP16X3 operands carry 64 x 64
    uint32_t lt = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++lt) {
      const int m0 = (tile / gn_tiles) * TILE_M + (int)cta_rank * kBM, n0 = (tile % num_n_tiles) * BN;
      const int gi = (tile % gn_tiles) / num_n_tiles;
      const int up_py = gi == 0 ? e.up_py : gp.up_py[gi - 1], up_px = gi == 0 ? e.up_px : gp.up_px[gi - 1];
      const int as = lt % ACC_STAGES;
      const uint32_t aph = (lt / ACC_STAGES) & 1;
      const int row_base = m0 + q * 32;
      // the 8 row slots this lane stores (row_base + i * 4 + rsub): element offset of the output row
      // (and of the residual row), or -1 for rows that produce nothing.  One division per tile, then
      // the map position advances incrementally (4 rows per slot).
      int64_t doff[8], roff[8];
      if constexpr (OUT != PP_OUT_PLANES) {
        const int m_first = row_base + rsub;
        if (e.up_hin == 0 && e.in_pad == 0 && e.out_pad == 0) {
          int rm = e.res_mod > 0 ? m_first % e.res_mod : 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m_first + 4 * i;
            doff[i] = m < e.m ? (int64_t)m * e.ldd : -1;
            roff[i] = (int64_t)(e.res_mod > 0 ? rm : m) * e.ldd;
            rm += 4;
            while (e.res_mod > 0 && rm >= e.res_mod) rm -= e.res_mod;
          }
        } else {
          const PadGeom ig = pad_geom(e.in_pad), og = pad_geom(e.out_pad);
          const int wp = e.in_w + ig.ex, hp = e.in_h + ig.ex;
          int pj = m_first % wp, t = m_first / wp;
          int pi = t % hp, pb = t / hp;
          const int ow = (e.up_hin ? 2 : 1) * e.in_w + og.ex, oh = (e.up_hin ? 2 : 1) * e.in_h + og.ex;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = (m_first + 4 * i < e.m) && pj >= ig.xo && pj < e.in_w + ig.xo && pi >= ig.yo && pi < e.in_h + ig.yo;
            const int ii = pi - ig.yo, jj = pj - ig.xo;
            const int oy = e.up_hin ? 2 * ii + up_py + og.yo : ii + og.yo, ox = e.up_hin ? 2 * jj + up_px + og.xo : jj + og.xo;
            const int64_t orow = (int64_t)(pb * oh + oy) * ow + ox;
            doff[i] = ok ? orow * e.ldd : -1;
            roff[i] = orow * e.ldd;
            pj += 4;
            while (pj >= wp) { pj -= wp; ++pi; }
            while (pi >= hp) { pi -= hp; ++pb; }
          }
        }
      }
      // short-K GEMMs with an fp32 residual (proj, patch embedding) are bound by the residual stream, which comes from
      // HBM: pull this tile's residual rows into L2 while its MMAs are still running (measured: proj 30.6 -> 29.0 us;
      // long-K fc2 76.5 -> 77.8 us, so not there)
      if constexpr (OUT == PP_OUT_F32 && ACCS == 1) {
        if (e.residual != nullptr && vec_ok && K <= 768) {
#pragma unroll 1
          for (int ch = half; ch < BN / 32; ch += 2) {
            const int col = n0 + ch * 32 + c4 * 4;
            if (col + 4 <= e.n) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (doff[i] >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(e.residual + roff[i] + col));
            }
          }
        }
      }
      ptx::mbar_wait(&tmem_full_bar[as], aph);
      ptx::tcgen05_fence_after();
      const uint32_t tacc = tmem_base + as * Cfg::ACC_COLS + lane_off;
#pragma unroll 1
      for (int ch = half; ch < BN / 32; ch += 2) {
        const int c0 = ch * 32;
        float v[32];
        ptx::tmem_ld_32x32b_x32(tacc + c0, v);
        if constexpr (ACCS == 2) {
          float w[32];
          ptx::tmem_ld_32x32b_x32(tacc + BN + c0, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += w[i];
        }
        if (ch + 2 >= BN / 32) {  // last TMEM read of this tile: hand the accumulator stage back
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) { if constexpr (PAIR) ptx::mbar_arrive_leader(&tmem_empty_bar[as]); else ptx::mbar_arrive(&tmem_empty_bar[as]); }
        }
        if (n0 + c0 >= e.n) continue;
        if constexpr (OUT == PP_OUT_PLANES) {
          // (M / plane, N, plane): lanes are consecutive pixels of one plane -> already coalesced
          const int row = row_base + lane;
          if (row < e.m) {
            const int64_t img = row / e.plane, pix = row % e.plane;
            float* dbase = reinterpret_cast<float*>(e.d) + (img * e.n + n0 + c0) * e.plane + pix;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int n = n0 + c0 + c;
              if (n < e.n) {
                const float sc = (e.scale ? __ldg(e.scale + n) : 1.f) * acc_scale, sh = e.shift ? __ldg(e.shift + n) : 0.f;
                float x = fmaf(v[c], sc, sh);
                if (e.act == PP_ACT_GELU) x = gelu_fast<1>(x);
                if (e.act == PP_ACT_RELU) x = fmaxf(x, 0.f);
                dbase[(int64_t)c * e.plane] = x;
              }
            }
          }
        } else {
          // transpose through swizzled shared memory: lane (row) writes 8 float4, then each lane
          // reads float4 column c4 of rows i * 4 + rsub -> 128-byte row segments per 8 lanes
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_f4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          __syncwarp();
          const int col = n0 + c0 + c4 * 4;
          if (col + 4 <= e.n && vec_ok) {
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.scale) sc = __ldg(reinterpret_cast<const float4*>(e.scale + col));
            if (e.shift) sh = __ldg(reinterpret_cast<const float4*>(e.shift + col));
            // FP16X3 operand output: the 64x operand scale is folded into the affine step, or (GELU,
            // which does not commute with scaling) into the activation itself
            constexpr bool kToUnits = OUT == PP_OUT_OPERAND && PREC == PP_PREC_FP16X3;
            const float osc = (kToUnits && e.act != PP_ACT_GELU) ? kOpScale : 1.0f;
            const float asc = acc_scale * osc;
            sc.x *= asc; sc.y *= asc; sc.z *= asc; sc.w *= asc;
            sh.x *= osc; sh.y *= osc; sh.z *= osc; sh.w *= osc;
            float4 x[8], rr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + rsub;
              x[i] = ld_shared_f4(stg + r * 128 + ((c4 ^ (r & 7)) << 4));
            }
            if constexpr (OUT == PP_OUT_F32) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e.residual != nullptr && doff[i] >= 0) rr[i] = *reinterpret_cast<const float4*>(e.residual + roff[i] + col);
              }
            }
            if (e.act == PP_ACT_GELU) {  // affine step + GELU on column pairs (FFMA2)
              constexpr int GS = kToUnits ? (int)kOpScale : 1;
              const uint64_t sc01 = pk2(sc.x, sc.y), sc23 = pk2(sc.z, sc.w), sh01 = pk2(sh.x, sh.y), sh23 = pk2(sh.z, sh.w);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                affine_gelu2<GS>(x[i].x, x[i].y, sc01, sh01);
                affine_gelu2<GS>(x[i].z, x[i].w, sc23, sh23);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                x[i].x = fmaf(x[i].x, sc.x, sh.x); x[i].y = fmaf(x[i].y, sc.y, sh.y);
                x[i].z = fmaf(x[i].z, sc.z, sh.z); x[i].w = fmaf(x[i].w, sc.w, sh.w);
              }
            }
            if (e.act == PP_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { x[i].x = fmaxf(x[i].x, 0.f); x[i].y = fmaxf(x[i].y, 0.f); x[i].z = fmaxf(x[i].z, 0.f); x[i].w = fmaxf(x[i].w, 0.f); }
            }
            if constexpr (OUT == PP_OUT_F32) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                x[i].x += rr[i].x; x[i].y += rr[i].y; x[i].z += rr[i].z; x[i].w += rr[i].w;
                if (doff[i] >= 0) *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.d) + doff[i] + col) = x[i];
              }
            } else {  // PP_OUT_OPERAND: doff = row * ldd elements of the logical operand
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (doff[i] >= 0) store_operand4_row<PREC>(e.d, doff[i], col, e.ldd, x[i]);
            }
          } else if (col < e.n) {  // ragged right edge or unaligned rows: scalar, rolled
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + rsub;
              bool ok;
              EpiParams eg = e;  // this group's phase
              eg.up_py = up_py; eg.up_px = up_px;
              const int64_t orow = map_out_row(eg, row_base + r, ok);
              if (!ok) continue;
              const int64_t rrow = e.res_mod > 0 ? (int)orow % e.res_mod : orow;
              const uint32_t src = stg + r * 128 + ((c4 ^ (r & 7)) << 4);
#pragma unroll 1
              for (int c = 0; c < 4 && col + c < e.n; ++c) {
                const float sc = (e.scale ? __ldg(e.scale + col + c) : 1.f) * acc_scale, sh = e.shift ? __ldg(e.shift + col + c) : 0.f;
                float x = fmaf(ld_shared_f1(src + 4 * c), sc, sh);
                if (e.act == PP_ACT_GELU) x = gelu_fast<1>(x);
                if (e.act == PP_ACT_RELU) x = fmaxf(x, 0.f);
                if constexpr (OUT == PP_OUT_F32) {
                  if (e.residual) x += e.residual[rrow * e.ldd + col + c];
                  reinterpret_cast<float*>(e.d)[orow * e.ldd + col + c] = x;
                } else {
                  store_operand<PREC>(e.d, orow, col + c, e.ldd, x);
                }
              }
            }
          }
          __syncwarp();  // staging tile is rewritten by the next chunk
        }
      }
      if (half >= BN / 32) {  // this warp had no chunk (BN == 32): still release the stage
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (PAIR) ptx::mbar_arrive_leader(&tmem_empty_bar[as]); else ptx::mbar_arrive(&tmem_empty_bar[as]); }
      }
    }
  }

  // pair: neither CTA may leave while the other can still signal its barriers or read its operands
  ptx::tcgen05_fence_before();
  if constexpr (PAIR) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tcgen05_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS); else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---- host side ----------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// K-major 16-bit operand (rows x row_elems), box = box_cols (64 / 32) elements x box_rows,
// 128- / 64-byte swizzle, OOB -> 0.
int make_operand_map(CUtensorMap* out, const void* base, int64_t rows, int64_t row_elems, int box_rows, int box_cols,
                     bool bf16) {
  typedef std::tuple<const void*, int64_t, int64_t, int, int, bool> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  const Key key(base, rows, row_elems, box_rows, box_cols, bf16);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return PP_OK; }
  }
  EncodeTiledFn enc = get_encode_fn();
  PP_REQUIRE(enc != nullptr, PP_ERR_CUDA, "cuTensorMapEncodeTiled not available from the CUDA driver");
  PP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, PP_ERR_INVALID, "GEMM operand %p not 16-byte aligned", base);
  const cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_elems * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PP_REQUIRE(r == CUDA_SUCCESS, PP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld elems=%lld box_rows=%d", (int)r,
             (long long)rows, (long long)row_elems, box_rows);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return PP_OK;
}

static int num_sms() { return device_sm_count(); }

// The group of the launch being dispatched (gemm_tc_launch_group sets it around gemm_tc_launch): the W operands, tap
// shifts and phases of groups 1.. (group 0 travels in the plain arguments).
struct GroupHost {
  int groups;
  const void* w[kMaxGroups - 1];
  int shift[kMaxGroups - 1][kMaxTaps];
  int up_py[kMaxGroups - 1], up_px[kMaxGroups - 1];
};
static thread_local const GroupHost* g_group = nullptr;

template <int BN, int SPLIT, bool BF16, int OUT, int ACCS, bool PAIR>
static int launch_tc(const pp_gemm_args& a, const EpiParams& e, const TapParams& tp, cudaStream_t st) {
  using Cfg = GemmCfg<BN, SPLIT, ACCS, PAIR>;
  static PerDeviceOnce attr_set;
  auto kern = gemm_tc_kernel<BN, SPLIT, BF16, OUT, ACCS, PAIR>;
  if (attr_set.first()) PP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int nops = SPLIT == 3 ? 2 : 1;
  CUtensorMap tma, tmw;
  int rc = make_operand_map(&tma, a.a, a.m, (int64_t)nops * tp.tap_k, kBM, Cfg::BK, BF16);
  if (rc) return rc;
  rc = make_operand_map(&tmw, a.w, a.n, (int64_t)nops * a.k, Cfg::BN_CTA, Cfg::BK, BF16);
  if (rc) return rc;
  GroupParams gp = {};
  gp.groups = 1;
  if (g_group != nullptr) {
    gp.groups = g_group->groups;
    for (int g = 1; g < gp.groups; ++g) {
      rc = make_operand_map(&gp.tm_w[g - 1], g_group->w[g - 1], a.n, (int64_t)nops * a.k, Cfg::BN_CTA, Cfg::BK, BF16);
      if (rc) return rc;
      for (int t = 0; t < kMaxTaps; ++t) gp.shift[g - 1][t] = g_group->shift[g - 1][t];
      gp.up_py[g - 1] = g_group->up_py[g - 1];
      gp.up_px[g - 1] = g_group->up_px[g - 1];
    }
  }
  constexpr int TILE_M = PAIR ? 2 * kBM : kBM;
  const int mt = (a.m + TILE_M - 1) / TILE_M, nt = (a.n + BN - 1) / BN;
  const int64_t tiles = (int64_t)mt * nt * gp.groups;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int na = L2Window::attach(attr, 0);
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (PAIR) {  // one cluster of 2 CTAs (one TPC) per 256-row tile, as many clusters as fit the SMs
    const int64_t pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  } else {
    cfg.gridDim = dim3((unsigned)(tiles < num_sms() ? tiles : num_sms()));
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const int k = a.k;
  PP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tma, tmw, k, mt, nt, tp, e, gp));
  count_launch();
  return PP_OK;
}

template <int BN, int SPLIT, bool BF16, int ACCS, bool PAIR>
static int launch_tc_out(const pp_gemm_args& a, const EpiParams& e, const TapParams& tp, cudaStream_t st) {
  switch (a.out_kind) {
    case PP_OUT_F32: return launch_tc<BN, SPLIT, BF16, PP_OUT_F32, ACCS, PAIR>(a, e, tp, st);
    case PP_OUT_OPERAND: return launch_tc<BN, SPLIT, BF16, PP_OUT_OPERAND, ACCS, PAIR>(a, e, tp, st);
    case PP_OUT_PLANES:
      if constexpr (BN <= 64 && !PAIR) return launch_tc<BN, SPLIT, BF16, PP_OUT_PLANES, ACCS, false>(a, e, tp, st);
  }
  set_error("pp_gemm: out_kind %d not built for tile_n %d", a.out_kind, BN);
  return PP_ERR_INVALID;
}

template <int SPLIT, bool BF16>
static int launch_tc_bn(int bn, int accs, bool pair, const pp_gemm_args& a, const EpiParams& e, const TapParams& tp,
                        cudaStream_t st) {
  if (a.out_kind == PP_OUT_PLANES && bn > 64) bn = 64;  // channel-plane output is built for narrow tiles (N = 17 logits)
  if (bn < 128 || a.out_kind == PP_OUT_PLANES) pair = false;  // CTA pairs are built for the wide tiles
  if constexpr (SPLIT == 3) {
    if (accs == 2) {  // two accumulators: 2 x BN columns per TMEM stage, so at most 128 wide
      switch (bn) {
        case 32: return launch_tc_out<32, SPLIT, BF16, 2, false>(a, e, tp, st);
        case 64: return launch_tc_out<64, SPLIT, BF16, 2, false>(a, e, tp, st);
        case 128: return pair ? launch_tc_out<128, SPLIT, BF16, 2, true>(a, e, tp, st)
                              : launch_tc_out<128, SPLIT, BF16, 2, false>(a, e, tp, st);
      }
    }
  }
  switch (bn) {
    case 32: return launch_tc_out<32, SPLIT, BF16, 1, false>(a, e, tp, st);
    case 64: return launch_tc_out<64, SPLIT, BF16, 1, false>(a, e, tp, st);
    case 128: return pair ? launch_tc_out<128, SPLIT, BF16, 1, true>(a, e, tp, st) : launch_tc_out<128, SPLIT, BF16, 1, false>(a, e, tp, st);
    case 192: return pair ? launch_tc_out<192, SPLIT, BF16, 1, true>(a, e, tp, st) : launch_tc_out<192, SPLIT, BF16, 1, false>(a, e, tp, st);
    case 256: return pair ? launch_tc_out<256, SPLIT, BF16, 1, true>(a, e, tp, st) : launch_tc_out<256, SPLIT, BF16, 1, false>(a, e, tp, st);
  }
  set_error("pp_gemm: unsupported tile_n %d", bn);
  return PP_ERR_INVALID;
}

// FP16X3 long-K GEMMs (fc2, deconv phases, 3x3 convolutions) keep the cross terms in a second
// accumulator: the fp32 accumulator is rounded once per MMA, and K / 16 x 3 roundings of one big
// accumulator is the dominant error term beyond K ~ 768 (measured: 1.7e-6 relative at K = 384,
// 1.3e-5 at K = 3456 with one accumulator).  They are MMA-bound at 128-wide tiles anyway.
constexpr int kTwoAccMinK = 768;

static int pick_tile_n(int n, int k, int prec) {
  if (n <= 32) return 32;
  if (n <= 64) return 64;
  if (prec == PP_PREC_FP16X3 && k > kTwoAccMinK) return 128;
  // widest tile that divides n: less L2 -> shared-memory traffic per MAC (see the header comment)
  if (n % 256 == 0) return 256;
  if (n % 192 == 0) return 192;
  return 128;
}

int gemm_tc_launch(const pp_gemm_args& a, const EpiParams& e, int tile_n, cudaStream_t st) {
  constexpr int kBK = 64;
  PP_REQUIRE(a.k % kBK == 0 && a.k > 0, PP_ERR_INVALID, "pp_gemm: k=%d must be a positive multiple of %d", a.k, kBK);
  TapParams tp = {};
  tp.taps = a.a_taps > 1 ? a.a_taps : 1;
  PP_REQUIRE(tp.taps <= kMaxTaps && a.k % tp.taps == 0 && (a.k / tp.taps) % kBK == 0, PP_ERR_INVALID,
             "pp_gemm: %d taps need k=%d to split into multiples of %d", tp.taps, a.k, kBK);
  tp.tap_k = a.k / tp.taps;
  for (int t = 0; t < tp.taps; ++t) tp.shift[t] = a.a_taps > 1 ? a.a_tap_shift[t] : 0;
  int bn = tile_n > 0 ? tile_n : pick_tile_n(a.n, a.k, a.precision);
  if (tile_n <= 0 && a.out_kind != PP_OUT_PLANES) {
    // few rows (the pooled 4x4 / 2x2 stages of the scalar branches): narrower tiles spread the long
    // K loop over more SMs - with fewer tiles than SMs the time is K / 16 x 3 x BN / 2 cycles per tile
    const int64_t mt = (a.m + kBM - 1) / kBM;
    if (mt * ((a.n + bn - 1) / bn) < num_sms()) {
      // cost model: rounds over the SMs x tile time (an MMA narrower than ~48 columns is latency-bound)
      const int widths[3] = {128, 64, 32};
      int64_t best_cost = -1;
      for (int i = 0; i < 3; ++i) {
        if (widths[i] > bn) continue;
        const int64_t tiles = mt * ((a.n + widths[i] - 1) / widths[i]);
        const int64_t cost = ((tiles + num_sms() - 1) / num_sms()) * (widths[i] > 48 ? widths[i] : 48);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; bn = widths[i]; }
      }
    }
  }
  const int accs = (a.precision == PP_PREC_FP16X3 && a.k > kTwoAccMinK && bn <= 128) ? 2 : 1;
  // CTA pairs pay off once there are enough 256-row tiles to occupy the 74 TPCs
  const bool pair = a.cta_pair == 2 || (a.cta_pair == 0 && (int64_t)((a.m + 255) / 256) * ((a.n + bn - 1) / bn) >= num_sms() / 2);
  switch (a.precision) {
    case PP_PREC_FP16X3: return launch_tc_bn<3, false>(bn, accs, pair, a, e, tp, st);
    case PP_PREC_BF16: return launch_tc_bn<1, true>(bn, accs, pair, a, e, tp, st);
    case PP_PREC_FP16: return launch_tc_bn<1, false>(bn, accs, pair, a, e, tp, st);
  }
  set_error("pp_gemm: precision %d is not a tensor-core mode", a.precision);
  return PP_ERR_INVALID;
}

// `count` GEMMs that differ only in their W operand, tap shifts and ConvTranspose2d phase (up_py, up_px), as one launch.
int gemm_tc_launch_group(const pp_gemm_args* a, const EpiParams& e0, int count, cudaStream_t st) {
  PP_REQUIRE(count >= 1 && count <= kMaxGroups, PP_ERR_INVALID, "grouped GEMM: %d groups outside [1, %d]", count, kMaxGroups);
  GroupHost gh = {};
  gh.groups = count;
  for (int g = 1; g < count; ++g) {
    const pp_gemm_args& b = a[g];
    PP_REQUIRE(b.precision == a[0].precision && b.m == a[0].m && b.n == a[0].n && b.k == a[0].k && b.a == a[0].a && b.d == a[0].d &&
                   b.scale == a[0].scale && b.shift == a[0].shift && b.residual == a[0].residual && b.act == a[0].act &&
                   b.out_kind == a[0].out_kind && b.ldd == a[0].ldd && b.a_taps == a[0].a_taps && b.in_pad == a[0].in_pad &&
                   b.out_pad == a[0].out_pad && b.in_h == a[0].in_h && b.in_w == a[0].in_w && b.up_hin == a[0].up_hin &&
                   b.up_win == a[0].up_win && b.tile_n == a[0].tile_n && b.cta_pair == a[0].cta_pair,
               PP_ERR_INVALID, "grouped GEMM: group %d differs from group 0 in more than W, tap shifts and phase", g);
    gh.w[g - 1] = b.w;
    for (int t = 0; t < kMaxTaps; ++t) gh.shift[g - 1][t] = b.a_taps > 1 ? b.a_tap_shift[t] : 0;
    gh.up_py[g - 1] = b.up_py;
    gh.up_px[g - 1] = b.up_px;
  }
  g_group = &gh;
  const int rc = gemm_tc_launch(a[0], e0, a[0].tile_n, st);
  g_group = nullptr;
  return rc;
}

}  // namespace pp
