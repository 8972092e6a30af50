// Global multi-head self-attention over the 192 ViT tokens on the 5th-generation tensor cores.
//
// Reference semantics: mmpretrain 1.2.0 MultiheadAttention.forward (see attention.cu) =
//   softmax(Q K^T * d_h^-0.5) V per (image, head), Q / K / V = column blocks of the qkv GEMM output.
//
// PERSISTENT CTAs (8 warps), two per SM, each walking the (image, head) units blockIdx.x, blockIdx.x + gridDim.x, ...
// (a static schedule: the units cost the same, and a launch carries no device-global scheduler state, so concurrent
// launches on other streams and CUDA-graph replays cannot interfere); the
// next unit's Q / K tiles are requested as soon as the current unit's last S = Q K^T has completed and its V as soon
// as the last P V has, so every load after the first hides behind the softmax / P V / read-out of the unit before
// (single-buffered shared memory: 80 KB per CTA), and barrier set-up / TMEM allocation happen once per CTA.
// A query row = a TMEM lane is shared by two threads
// (warps w and w + 4 may access the same lane quarter) that split its 192 keys in halves:
//   * TMA (cp.async.bulk.tensor, 64- / 128-byte swizzle) stages Q (two 128-row tiles: rows 0-127 and
//     128-255, of which 128-191 are this image's), K and V straight out of the qkv operand,
//   * S = Q K^T: tcgen05.mma M = 128, N = 192, K = d_h, fp32 accumulator in TMEM (192 columns),
//   * softmax: tcgen05.ld 32 columns at a time, row max (halves combined through shared memory), exp2,
//     row sum; P goes back INTO THE SAME TMEM COLUMNS as packed 16-bit pairs
//     (tcgen05.st) - per 32-key chunk [hi : 16 columns | lo : 16 columns] in that chunk's own columns - no shared-memory round trip,
//   * O = P V: tcgen05.mma with the A operand read from TMEM and V consumed as stored (keys x d_h,
//     i.e. an MN-major B operand), 12 k-steps of 16 keys into a d_h-column TMEM accumulator,
//   * O / rowsum leaves as the proj GEMM's A operand.
// FP16X3 parity mode: the same 3-term split as the GEMMs (hi.hi + hi.lo + lo.hi of 64x-scaled
// operands, one fp32 accumulator); Q, K, V arrive pre-split from the qkv GEMM epilogue, P is split in
// registers (the factor 64 of P rides in the exponent).
// While one CTA of the SM is in its softmax (ALU / MUFU bound), the other one loads or runs its MMAs.
#include <cuda.h>

#include "engine_ops.cuh"
#include "ptx.cuh"

#include <math.h>

namespace pp {

int make_operand_map(CUtensorMap* out, const void* base, int64_t rows, int64_t row_elems, int box_rows, int box_cols,
                     bool bf16);  // gemm_tc.cu

namespace {

constexpr int kTcThreads = 256;
constexpr int kNTok = 192;

int num_sms_att() { return device_sm_count(); }

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Barrier wait of the 256 softmax threads: back off between polls so that the waiting CTA does not take
// issue slots from the CTA that shares the SM (bounded like ptx::mbar_wait: a lost arrival traps).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if (++spins > (1u << 22)) __trap();
  }
}

template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_st<32>(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// packed fp32 pairs (pk2 / upk2 / add2 / fma2): common.cuh
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&r)[8]) {  // p 32-byte aligned
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool BF16>
__device__ __forceinline__ uint32_t pack_pair(float x, float y) {  // x in the low half
  if constexpr (BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {  // operand units in, [hi | lo] out
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int DH, int SPLIT>
struct AttCfg {
  static constexpr int NOPS = SPLIT == 3 ? 2 : 1;
  static constexpr int ROWB = DH * 2;               // bytes per smem row = swizzle span (64 or 128)
  static constexpr int Q_BYTES = 128 * ROWB;        // one Q tile, one plane
  static constexpr int KV_BYTES = kNTok * ROWB;     // K or V, one plane
  static constexpr int OFF_K = 2 * NOPS * Q_BYTES;
  static constexpr int OFF_V = OFF_K + NOPS * KV_BYTES;
  static constexpr int OFF_BAR = OFF_V + NOPS * KV_BYTES;
  static constexpr int LOAD_QK_BYTES = NOPS * (2 * Q_BYTES + KV_BYTES);  // both Q tiles + K: free again after the unit's second S
  static constexpr int LOAD_V_BYTES = NOPS * KV_BYTES;                   // V: free again after the unit's second P V
  static constexpr int OFF_X = OFF_BAR + 64;        // row max / row sum halves: float [2][2][128]; next-unit mailbox at OFF_BAR + 48
  static constexpr int SMEM_BYTES = OFF_X + 2 * 2 * 128 * 4 + 1024;  // + alignment slack
  static constexpr int OCOLS = SPLIT == 3 ? 2 * DH : DH;  // FP16X3: [P V_hi + P_lo V_hi | P_hi V_lo], added in the read-out
  static constexpr int TMEM_COLS = 192 + OCOLS <= 256 ? 256 : 512;  // S / P: 192 columns, O at 192
  static_assert(DH == 32 || DH == 64, "head width");
  static_assert(192 + OCOLS <= TMEM_COLS, "accumulators exceed the allocation");
};

template <int DH, int SPLIT, bool BF16>
__global__ void __launch_bounds__(kTcThreads, DH == 32 ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const int heads,
                    const int units, uint16_t* __restrict__ out_op) {
  using Cfg = AttCfg<DH, SPLIT>;
  constexpr int NOPS = Cfg::NOPS, ROWB = Cfg::ROWB;
  extern __shared__ uint8_t att_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full_bar1 = full_bar + 1;
  uint64_t* s_bar = full_bar + 2;
  uint64_t* o_bar = full_bar + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full_bar + 4);
  float* x_max = reinterpret_cast<float*>(smem + Cfg::OFF_X);  // [half][row]
  float* x_sum = x_max + 2 * 128;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * DH;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_kv);
    ptx::mbar_init(full_bar, 1);
    ptx::mbar_init(full_bar1, 1);
    ptx::mbar_init(s_bar, 1);
    ptx::mbar_init(o_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // qkv comes from the qkv GEMM

  auto sQ = [&](int tile, int part) { return smem + (tile * NOPS + part) * Cfg::Q_BYTES; };
  auto sK = [&](int part) { return smem + Cfg::OFF_K + part * Cfg::KV_BYTES; };
  auto sV = [&](int part) { return smem + Cfg::OFF_V + part * Cfg::KV_BYTES; };

  auto load_qk = [&](int unit) {  // thread 0
    const int row0 = (unit / heads) * kNTok, hc = (unit % heads) * DH;
    ptx::mbar_arrive_expect_tx(full_bar, Cfg::LOAD_QK_BYTES);
#pragma unroll
    for (int part = 0; part < NOPS; ++part) {
      const int c = part * 3 * D + hc;  // q | k | v column blocks, lo plane 3 D further
      ptx::tma_load_2d(sQ(0, part), &tm_q, full_bar, c, row0);
      ptx::tma_load_2d(sK(part), &tm_kv, full_bar, c + D, row0);
    }
#pragma unroll
    for (int part = 0; part < NOPS; ++part)  // rows 192.. belong to the next image (or are zero-filled): never stored
      ptx::tma_load_2d(sQ(1, part), &tm_q, full_bar, part * 3 * D + hc, row0 + 128);
  };
  auto load_v = [&](int unit) {  // thread 0
    const int row0 = (unit / heads) * kNTok, hc = (unit % heads) * DH;
    ptx::mbar_arrive_expect_tx(full_bar1, Cfg::LOAD_V_BYTES);
#pragma unroll
    for (int part = 0; part < NOPS; ++part) ptx::tma_load_2d(sV(part), &tm_kv, full_bar1, part * 3 * D + hc + 2 * D, row0);
  };

  constexpr uint32_t idesc_s = ptx::make_idesc_f16(BF16, 128, kNTok);
  constexpr uint32_t idesc_o = ptx::make_idesc_f16(BF16, 128, DH) | (1u << 16);  // B (= V, keys x d_h) is MN-major
  const uint32_t t_s = tmem_base, t_o = tmem_base + 192;

  auto issue_s = [&](int tile) {  // thread 0: S = Q K^T for one query tile
    ptx::tcgen05_fence_after();
    const uint32_t qh = ptx::kmajor_desc_lo(ptx::smem_u32(sQ(tile, 0))), kh = ptx::kmajor_desc_lo(ptx::smem_u32(sK(0)));
    const uint32_t ql = qh + (Cfg::Q_BYTES >> 4), kl = kh + (Cfg::KV_BYTES >> 4);
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      const uint64_t da = ptx::kmajor_desc<ROWB>(qh + 2 * kk), db = ptx::kmajor_desc<ROWB>(kh + 2 * kk);
      ptx::umma_f16(t_s, da, db, idesc_s, kk != 0 ? 1u : 0u);
      if constexpr (SPLIT == 3) {
        ptx::umma_f16(t_s, da, ptx::kmajor_desc<ROWB>(kl + 2 * kk), idesc_s, 1u);
        ptx::umma_f16(t_s, ptx::kmajor_desc<ROWB>(ql + 2 * kk), db, idesc_s, 1u);
      }
    }
    ptx::umma_commit(s_bar);
  };
  auto issue_o = [&](uint32_t v_parity) {  // thread 0: O = P V, P read from TMEM
    ptx::mbar_wait(full_bar1, v_parity);  // V has landed
    ptx::tcgen05_fence_after();
    const uint32_t vh = ptx::kmajor_desc_lo(ptx::smem_u32(sV(0)));
    const uint32_t vl = vh + (Cfg::KV_BYTES >> 4);
    constexpr uint32_t KSTEP = (16 * ROWB) >> 4;  // 16 keys further
    constexpr int PCH = 32;                       // TMEM columns per 32-key chunk of P (its own S columns)
    // An MMA this small costs ~90 cycles whatever its N (measured: the tensor pipe's instruction rate bounds the kernel), so
    // the split product takes TWO instructions per 16-key step instead of three: P_hi x [V_hi | V_lo] as ONE N = 2 d_h
    // MMA - V is an MN-major operand whose leading-dimension byte offset (the stride between its column atoms) is set to
    // the distance between the hi and lo planes, so columns d_h.. of "B" are V_lo - plus P_lo x V_hi into the first d_h
    // columns.  The read-out adds the two column blocks.
    constexpr uint32_t idesc_o2 = ptx::make_idesc_f16(BF16, 128, 2 * DH) | (1u << 16);
    constexpr uint32_t KVLO = Cfg::KV_BYTES >> 4;
    (void)vl;
#pragma unroll
    for (int j = 0; j < kNTok / 16; ++j) {
      const uint32_t a_hi = t_s + PCH * (j >> 1) + 8 * (j & 1);
      const uint32_t lo_w = (vh + KSTEP * j) & 0x3fffu;
      const uint64_t db = ptx::kmajor_desc<ROWB>(lo_w | (1u << 16));
      if constexpr (SPLIT == 3) {
        umma_f16_ts(t_o, a_hi, ptx::kmajor_desc<ROWB>(lo_w | (KVLO << 16)), idesc_o2, j != 0 ? 1u : 0u);
        umma_f16_ts(t_o, a_hi + 16, db, idesc_o, 1u);
      } else {
        umma_f16_ts(t_o, a_hi, db, idesc_o, j != 0 ? 1u : 0u);
      }
    }
    ptx::umma_commit(o_bar);
  };

  int unit = blockIdx.x;  // static schedule: unit, unit + gridDim.x, ... (grid <= units)
  if (threadIdx.x == 0) {
    load_qk(unit);
    load_v(unit);
    ptx::mbar_wait(full_bar, 0);
    issue_s(0);
  }

  // exp2 argument scale: d_h^-0.5 * log2(e); FP16X3 scores carry the operand scale 64 * 64
  const float c_exp = rsqrtf((float)DH) * 1.4426950408889634f * (SPLIT == 3 ? kAccScaleInv : 1.0f);
  const float p_exp = SPLIT == 3 ? 6.0f : 0.0f;  // P leaves the exponential already in operand units (x 2^6)
  const int q = warp & 3, hf = warp >> 2;  // TMEM lane quarter, and which half of the keys / of the O columns
  const int row = q * 32 + lane;           // query row of the tile = TMEM lane
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  constexpr int CHH = kNTok / 64;          // 32-key chunks per half
  constexpr int PCH = 32;  // P of a 32-key chunk overwrites that chunk's own S columns (no thread reads another's)

#pragma unroll 1
  for (uint32_t it = 0; unit < units; ++it) {
  const int b = unit / heads, h = unit % heads;
  const int next = unit + (int)gridDim.x;
#pragma unroll 1
  for (int tile = 0; tile < 2; ++tile) {
    const bool active = tile == 0 || q < 2;  // tile 1: only rows 128..191 are real
    mbar_wait_relaxed(s_bar, tile);  // two S per unit: the barrier's phase parity is the tile index
    ptx::tcgen05_fence_after();
    if (threadIdx.x == 0 && tile == 1 && next < units) load_qk(next);  // both S of this unit are done: its Q and K tiles are free
    float mx = -INFINITY;
    if (active) {
#pragma unroll 1
      for (int ch = hf * CHH; ch < (hf + 1) * CHH; ++ch) {
        float v[32];
        ptx::tmem_ld_32x32b_x32(t_s + lane_base + 32 * ch, v);
        float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};  // four independent chains
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          m4[0] = fmaxf(m4[0], fmaxf(v[i], v[i + 1])); m4[1] = fmaxf(m4[1], fmaxf(v[i + 2], v[i + 3]));
          m4[2] = fmaxf(m4[2], fmaxf(v[i + 4], v[i + 5])); m4[3] = fmaxf(m4[3], fmaxf(v[i + 6], v[i + 7]));
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      x_max[hf * 128 + row] = mx;
      // the two halves of a row live in warps q and q + 4: only those 64 threads have to meet
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, x_max[(hf ^ 1) * 128 + row]);
      float l4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains (fixed order: deterministic)
#pragma unroll 1
      for (int ch = hf * CHH; ch < (hf + 1) * CHH; ++ch) {
        float v[32];
        ptx::tmem_ld_32x32b_x32(t_s + lane_base + 32 * ch, v);
        {  // exp2(((s - max) * c) + p) and the four row-sum chains on packed fp32 pairs (FFMA2 / FADD2: two lanes per
           // issue slot, same operations and order per element as the scalar form; (s - max) first: exact near the maximum)
          const uint64_t nmx2 = pk2(-mx, -mx), ce2 = pk2(c_exp, c_exp), pe2 = pk2(p_exp, p_exp);
          uint64_t la = pk2(l4[0], l4[1]), lb = pk2(l4[2], l4[3]);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a, b;
            upk2(fma2(add2(pk2(v[i], v[i + 1]), nmx2), ce2, pe2), a, b);
            v[i] = ex2f(a);
            v[i + 1] = ex2f(b);
            if ((i & 2) == 0) la = add2(la, pk2(v[i], v[i + 1]));
            else lb = add2(lb, pk2(v[i], v[i + 1]));
          }
          upk2(la, l4[0], l4[1]);
          upk2(lb, l4[2], l4[3]);
        }
        if constexpr (SPLIT == 3) {
          uint32_t r[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) split_pair(v[2 * i], v[2 * i + 1], r[i], r[16 + i]);
          tmem_st<32>(t_s + lane_base + PCH * ch, r);
        } else {
          uint32_t r[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = pack_pair<BF16>(v[2 * i], v[2 * i + 1]);
          tmem_st<16>(t_s + lane_base + PCH * ch, r);
        }
      }
      x_sum[hf * 128 + row] = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      tmem_st_wait();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();  // P of all rows is in TMEM, the row-sum halves in shared memory
    if (threadIdx.x == 0) issue_o(it & 1);
    mbar_wait_relaxed(o_bar, tile);
    ptx::tcgen05_fence_after();
    if (threadIdx.x == 0) {  // the next S overlaps the read-out of O below (disjoint columns)
      if (tile == 0) {
        issue_s(1);
      } else if (next < units) {  // both P V of this unit are done: V is free, and so are the S / P columns
        load_v(next);
        ptx::mbar_wait(full_bar, (it + 1) & 1);
        issue_s(0);
      }
    }
    if (active) {
      // FP16X3: O carries 64 (P) * 64 (V) and the row sum carries 64, so O / l is already in operand units
      const float inv = 1.0f / (x_sum[row] + x_sum[128 + row]);
      const size_t orow = (size_t)b * kNTok + tile * 128 + row;
      uint16_t* d = out_op + orow * (NOPS * D) + h * DH + hf * (DH / 2);  // this thread's half of the head's columns
#pragma unroll
      for (int c0 = 0; c0 < DH / 2; c0 += 16) {
        float o[16];
        tmem_ld_x16(t_o + lane_base + hf * (DH / 2) + c0, o);
        if constexpr (SPLIT == 3) {  // + the P_hi V_lo block
          float o2[16];
          tmem_ld_x16(t_o + lane_base + DH + hf * (DH / 2) + c0, o2);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] += o2[i];
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if constexpr (SPLIT == 3) split_pair(o[2 * i] * inv, o[2 * i + 1] * inv, hi[i], lo[i]);
          else hi[i] = pack_pair<BF16>(o[2 * i] * inv, o[2 * i + 1] * inv);
        }
        // 16 columns = one 32-byte sector per plane: a single 256-bit store each (sm_100 STG.256) instead of two
        // 128-bit stores that each leave the lane's sector half written
        st_global_v8(d + c0, hi);
        if constexpr (SPLIT == 3) st_global_v8(d + D + c0, lo);
      }
    }
    ptx::tcgen05_fence_before();
    __syncthreads();  // O has been read: the next tile's P V may overwrite it (and the exchange arrays are free)
  }
  unit = next;
  }
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int DH, int SPLIT, bool BF16>
int launch_tc(const void* qkv_op, int batch, int heads, void* out_op, cudaStream_t st) {
  using Cfg = AttCfg<DH, SPLIT>;
  auto kern = attention_tc_kernel<DH, SPLIT, BF16>;
  static PerDeviceOnce attr_set;
  if (attr_set.first()) PP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int64_t rows = (int64_t)batch * kNTok, row_elems = (int64_t)Cfg::NOPS * 3 * heads * DH;
  CUtensorMap tq, tkv;
  int rc = make_operand_map(&tq, qkv_op, rows, row_elems, 128, DH, BF16);
  if (rc) return rc;
  rc = make_operand_map(&tkv, qkv_op, rows, row_elems, kNTok, DH, BF16);
  if (rc) return rc;
  const int units = batch * heads;
  const int resident = (DH == 32 ? 2 : 1) * num_sms_att();
  PP_CHECK_CUDA(launch_pdl(kern, dim3(units < resident ? units : resident), dim3(kTcThreads), Cfg::SMEM_BYTES, st, tq, tkv,
                           heads, units, reinterpret_cast<uint16_t*>(out_op)));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

template <int DH>
int launch_tc_prec(int prec, const void* qkv_op, int batch, int heads, void* out_op, cudaStream_t st) {
  switch (prec) {
    case PP_PREC_FP16X3: return launch_tc<DH, 3, false>(qkv_op, batch, heads, out_op, st);
    case PP_PREC_BF16: return launch_tc<DH, 1, true>(qkv_op, batch, heads, out_op, st);
    case PP_PREC_FP16: return launch_tc<DH, 1, false>(qkv_op, batch, heads, out_op, st);
  }
  set_error("attention: precision %d is not a tensor-core mode", prec);
  return PP_ERR_INVALID;
}

}  // namespace

// Same contract as launch_attention_mma (attention.cu).
int launch_attention_tc(int prec, const void* qkv_op, int batch, int n, int heads, int dh, void* out_op, cudaStream_t st) {
  PP_REQUIRE(attention_mma_supported(n, dh), PP_ERR_UNSUPPORTED,
             "tensor-core attention is built for 192 tokens and head width 32 / 64 (got %d tokens, width %d)", n, dh);
  if (batch == 0) return PP_OK;
  PP_REQUIRE((reinterpret_cast<uintptr_t>(out_op) & 31) == 0, PP_ERR_INVALID, "attention: out_op %p must be 32-byte aligned", out_op);
  return dh == 32 ? launch_tc_prec<32>(prec, qkv_op, batch, heads, out_op, st)
                  : launch_tc_prec<64>(prec, qkv_op, batch, heads, out_op, st);
}

}  // namespace pp
