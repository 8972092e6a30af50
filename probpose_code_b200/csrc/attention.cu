// Global multi-head self-attention over the 192 ViT tokens on the tensor cores.
//
// Reference semantics: mmpretrain 1.2.0 MultiheadAttention.forward =
//   qkv = Linear(x).reshape(B, N, 3, heads, d_h).permute(2, 0, 3, 1, 4)
//   x   = F.scaled_dot_product_attention(q, k, v)          (scale d_h^-0.5, no mask, no dropout)
//   x   = x.transpose(1, 2).reshape(B, N, D)
// (SURVEY.md section 8c "Backbone"; config td-pm_ProbPose-small_8xb64-210e_coco-256x192.py:56-67).
//
// One CTA per (image, head), 6 warps; a warp owns 16 query rows at a time and the WHOLE 192-key
// score row lives in its registers, so softmax needs no online rescaling:
//   S = Q K^T   mma.sync m16n8k16, K fragments by ldmatrix from padded shared memory, 96 keys at
//               a time (the score chunk lives in registers; running max / sum between the chunks)
//   P = exp2((S - rowmax) * scale * log2 e)                 fp32, row sums by quad shuffles
//   O = P V     the S accumulator layout IS the A-fragment layout of the second product
//   out = O / rowsum  -> written as the proj GEMM's A operand
// In the FP16X3 parity mode every product is the same 3-term split the GEMMs use (hi.hi + hi.lo +
// lo.hi of 64x-scaled operands, one fp32 accumulator), with Q, K, V arriving pre-split from the qkv
// GEMM epilogue and P split in registers (the factor 64 of P is folded into the exponent).
//
// Attention is 7.7 % of the block FLOPs with d_h = 32 (K = 32 per QK^T product) and is bound by the
// exp / conversion work, not by tensor throughput, so it uses the register-level mma.sync path; the
// tcgen05 pipeline is spent on the GEMMs (gemm_tc.cu).
#include "engine_ops.cuh"

#include <math.h>

namespace pp {

namespace {

constexpr int kAttThreads = 192;

template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float ex2_fast(float x) {  // 2^x, MUFU.EX2 (flushes denormal results: irrelevant after max-subtraction)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Two fp32 -> packed 16-bit pair (x in the low half), and the scaled residual pair for FP16X3.
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float x, float y) {
  if constexpr (BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
// x, y are already in operand units (64x the value they stand for)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// NTOK tokens, DH head width, SPLIT = 3 (FP16X3) or 1, BF16 element type for SPLIT == 1.
template <int NTOK, int DH, int SPLIT, bool BF16>
__global__ void __launch_bounds__(kAttThreads, DH == 32 ? 3 : 1)
attention_mma_kernel(const uint16_t* __restrict__ qkv_op, int heads, uint16_t* __restrict__ out_op) {
  constexpr int NOPS = SPLIT == 3 ? 2 : 1;
  constexpr int KT = NTOK / 8;        // key tiles of 8
  constexpr int RT = NTOK / 16;       // query row tiles of 16
  constexpr int KS = DH / 16;         // k-steps of the QK^T product
  constexpr int DT = DH / 8;          // output column tiles of the PV product
  constexpr int ROWB = DH * 2 + 16;   // padded shared-memory row (bytes): conflict-free ldmatrix
  constexpr int ARR = NTOK * ROWB;    // one K or V plane
  static_assert(NTOK % 16 == 0 && DH % 16 == 0, "tile shapes");

  extern __shared__ __align__(16) uint8_t att_smem[];  // [K hi | K lo | V hi | V lo]
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int D = heads * DH;
  const int RS = NOPS * 3 * D;  // qkv operand row stride (elements)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint16_t* base = qkv_op + (size_t)b * NTOK * RS + h * DH;
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(att_smem);

  // ---- K, V (hi and lo planes) -> shared memory, 16-byte cp.async chunks ----
  {
    constexpr int CPR = DH / 8;  // chunks per row per plane (a power of two)
#pragma unroll
    for (int pl = 0; pl < 2 * NOPS; ++pl) {
      const int kv = pl / NOPS, part = pl % NOPS;  // kv: 0 = K, 1 = V; part: 0 = hi, 1 = lo
      const uint16_t* src0 = base + (1 + kv) * D + part * 3 * D;
      const uint32_t dst0 = s_base + pl * ARR;
      for (int i = threadIdx.x; i < NTOK * CPR; i += kAttThreads) {
        const int r = i / CPR, c = i % CPR;
        cp_async16(dst0 + r * ROWB + c * 16, src0 + (size_t)r * RS + c * 8);
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }
  const uint32_t sK[2] = {s_base, s_base + ARR};
  const uint32_t sV[2] = {s_base + NOPS * ARR, s_base + (NOPS + 1) * ARR};

  const int g = lane >> 2, t4 = lane & 3;
  // exp2 argument scale: d_h^-0.5 * log2(e); FP16X3 scores carry the operand scale 64 * 64
  const float c_exp = rsqrtf((float)DH) * 1.4426950408889634f * (SPLIT == 3 ? kAccScaleInv : 1.0f);
  const float p_exp = SPLIT == 3 ? 6.0f : 0.0f;  // P leaves the exponential already in operand units (x 2^6)

  for (int rt = warp; rt < RT; rt += kAttThreads / 32) {
    // ---- Q fragments (A operand, row-major): rows rt*16 + g (+8), k = ks*16 + 2 t4 (+8) ----
    uint32_t qh[KS][4], ql[KS][4];
    {
      const uint16_t* q0 = base + (size_t)(rt * 16 + g) * RS + 2 * t4;
      const uint16_t* q1 = q0 + (size_t)8 * RS;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qh[ks][0] = *reinterpret_cast<const uint32_t*>(q0 + ks * 16);
        qh[ks][1] = *reinterpret_cast<const uint32_t*>(q1 + ks * 16);
        qh[ks][2] = *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8);
        qh[ks][3] = *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8);
        if constexpr (SPLIT == 3) {
          ql[ks][0] = *reinterpret_cast<const uint32_t*>(q0 + 3 * D + ks * 16);
          ql[ks][1] = *reinterpret_cast<const uint32_t*>(q1 + 3 * D + ks * 16);
          ql[ks][2] = *reinterpret_cast<const uint32_t*>(q0 + 3 * D + ks * 16 + 8);
          ql[ks][3] = *reinterpret_cast<const uint32_t*>(q1 + 3 * D + ks * 16 + 8);
        }
      }
    }

    // ---- keys in chunks of KC: S = Q K^T (registers) -> running max / sum -> O += P V ----
    // (flash-style rescaling between the chunks; exact softmax up to fp32 rounding)
    constexpr int KC = 48, KTC = KC / 8, NCH = NTOK / KC;
    static_assert(NTOK % KC == 0 && KTC % 2 == 0, "key chunking");
    float o0[DT][4];
#pragma unroll
    for (int dt = 0; dt < DT; ++dt)
#pragma unroll
      for (int i = 0; i < 4; ++i) o0[dt][i] = 0.f;
    float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    // ldmatrix lane address: matrix (lane / 8) = 16-byte dh chunk, row = key (lane % 8)
    const uint32_t k_lane = (uint32_t)((lane & 7) * ROWB + (lane >> 3) * 16);
    // ldmatrix.trans lane address: matrices {keys 0-7, keys 8-15} x {dh chunk c, c + 1}
    const uint32_t v_lane = (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * ROWB + (lane >> 4) * 16);
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
      float s[KTC][4];
#pragma unroll
      for (int nt = 0; nt < KTC; ++nt) {
        float a0[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t krow = (uint32_t)((ch * KTC + nt) * 8 * ROWB);
#pragma unroll
        for (int kp = 0; kp < KS / 2; ++kp) {  // one ldmatrix.x4 covers two k-steps (32 dh)
          uint32_t kh[4], kl[4];
          ldmatrix_x4(kh, sK[0] + krow + kp * 64 + k_lane);
          if constexpr (SPLIT == 3) ldmatrix_x4(kl, sK[1] + krow + kp * 64 + k_lane);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int ks = kp * 2 + u;
            mma16816<BF16>(a0, qh[ks], kh[2 * u], kh[2 * u + 1]);
            if constexpr (SPLIT == 3) {
              mma16816<BF16>(a0, qh[ks], kl[2 * u], kl[2 * u + 1]);
              mma16816<BF16>(a0, ql[ks], kh[2 * u], kh[2 * u + 1]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s[nt][i] = a0[i];
      }
      // running max over the rows g and g + 8 of the tile
      float c0 = mx0, c1 = mx1;
#pragma unroll
      for (int nt = 0; nt < KTC; ++nt) {
        c0 = fmaxf(c0, fmaxf(s[nt][0], s[nt][1]));
        c1 = fmaxf(c1, fmaxf(s[nt][2], s[nt][3]));
      }
      c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1)); c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
      c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
      const float r0 = ex2_fast((mx0 - c0) * c_exp), r1 = ex2_fast((mx1 - c1) * c_exp);  // 2^-inf = 0 on the first chunk
      mx0 = c0; mx1 = c1;
      l0 *= r0; l1 *= r1;
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) { o0[dt][0] *= r0; o0[dt][1] *= r0; o0[dt][2] *= r1; o0[dt][3] *= r1; }
#pragma unroll
      for (int nt = 0; nt < KTC; ++nt) {
        // (s - max) first: exact for scores near the maximum, where the probabilities matter
        s[nt][0] = ex2_fast(fmaf(s[nt][0] - mx0, c_exp, p_exp)); s[nt][1] = ex2_fast(fmaf(s[nt][1] - mx0, c_exp, p_exp));
        s[nt][2] = ex2_fast(fmaf(s[nt][2] - mx1, c_exp, p_exp)); s[nt][3] = ex2_fast(fmaf(s[nt][3] - mx1, c_exp, p_exp));
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      // O += P V for this chunk's keys
#pragma unroll
      for (int j = 0; j < KTC / 2; ++j) {  // k-steps of 16 keys
        uint32_t ph[4], pl[4];
        if constexpr (SPLIT == 3) {
          split2(s[2 * j][0], s[2 * j][1], ph[0], pl[0]);
          split2(s[2 * j][2], s[2 * j][3], ph[1], pl[1]);
          split2(s[2 * j + 1][0], s[2 * j + 1][1], ph[2], pl[2]);
          split2(s[2 * j + 1][2], s[2 * j + 1][3], ph[3], pl[3]);
        } else {
          ph[0] = pack2<BF16>(s[2 * j][0], s[2 * j][1]);
          ph[1] = pack2<BF16>(s[2 * j][2], s[2 * j][3]);
          ph[2] = pack2<BF16>(s[2 * j + 1][0], s[2 * j + 1][1]);
          ph[3] = pack2<BF16>(s[2 * j + 1][2], s[2 * j + 1][3]);
        }
        const uint32_t vrow = (uint32_t)((ch * KC + j * 16) * ROWB);
#pragma unroll
        for (int dp = 0; dp < DT / 2; ++dp) {  // one ldmatrix.x4.trans covers two 8-wide dh tiles
          uint32_t vh[4], vl[4];
          ldmatrix_x4_trans(vh, sV[0] + vrow + dp * 32 + v_lane);
          if constexpr (SPLIT == 3) ldmatrix_x4_trans(vl, sV[1] + vrow + dp * 32 + v_lane);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int dt = dp * 2 + u;
            mma16816<BF16>(o0[dt], ph, vh[2 * u], vh[2 * u + 1]);
            if constexpr (SPLIT == 3) {
              mma16816<BF16>(o0[dt], ph, vl[2 * u], vl[2 * u + 1]);
              mma16816<BF16>(o0[dt], pl, vh[2 * u], vh[2 * u + 1]);
            }
          }
        }
      }
    }
    // the per-lane partial row sums (4 lanes share a row)
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- normalise and write the proj GEMM's A operand (rows b*NTOK + ..., cols h*DH + ...) ----
    // FP16X3: O carries 64 (P) * 64 (V) and the row sum carries 64, so O / l is already in operand units
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const size_t orow0 = (size_t)b * NTOK + rt * 16 + g;
    uint16_t* d0 = out_op + orow0 * (NOPS * D) + h * DH + 2 * t4;
    uint16_t* d1 = d0 + (size_t)8 * (NOPS * D);
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const float x0 = o0[dt][0] * inv0, x1 = o0[dt][1] * inv0, y0 = o0[dt][2] * inv1, y1 = o0[dt][3] * inv1;
      if constexpr (SPLIT == 3) {
        uint32_t hi, lo;
        split2(x0, x1, hi, lo);
        *reinterpret_cast<uint32_t*>(d0 + dt * 8) = hi;
        *reinterpret_cast<uint32_t*>(d0 + D + dt * 8) = lo;
        split2(y0, y1, hi, lo);
        *reinterpret_cast<uint32_t*>(d1 + dt * 8) = hi;
        *reinterpret_cast<uint32_t*>(d1 + D + dt * 8) = lo;
      } else {
        *reinterpret_cast<uint32_t*>(d0 + dt * 8) = pack2<BF16>(x0, x1);
        *reinterpret_cast<uint32_t*>(d1 + dt * 8) = pack2<BF16>(y0, y1);
      }
    }
  }
}

template <int NTOK, int DH, int SPLIT, bool BF16>
int launch_mma(const void* qkv_op, int batch, int heads, void* out_op, cudaStream_t st) {
  constexpr int NOPS = SPLIT == 3 ? 2 : 1;
  constexpr int SMEM = 2 * NOPS * NTOK * (DH * 2 + 16);
  auto kern = attention_mma_kernel<NTOK, DH, SPLIT, BF16>;
  static PerDeviceOnce attr_set;
  if (attr_set.first()) PP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  PP_CHECK_CUDA(launch_pdl(kern, dim3(batch * heads), dim3(kAttThreads), SMEM, st, reinterpret_cast<const uint16_t*>(qkv_op), heads,
                           reinterpret_cast<uint16_t*>(out_op)));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

template <int DH>
int launch_mma_prec(int prec, const void* qkv_op, int batch, int heads, void* out_op, cudaStream_t st) {
  switch (prec) {
    case PP_PREC_FP16X3: return launch_mma<192, DH, 3, false>(qkv_op, batch, heads, out_op, st);
    case PP_PREC_BF16: return launch_mma<192, DH, 1, true>(qkv_op, batch, heads, out_op, st);
    case PP_PREC_FP16: return launch_mma<192, DH, 1, false>(qkv_op, batch, heads, out_op, st);
  }
  set_error("attention: precision %d is not a tensor-core mode", prec);
  return PP_ERR_INVALID;
}

}  // namespace

bool attention_mma_supported(int n, int dh) { return n == 192 && (dh == 32 || dh == 64); }

// qkv_op: operand (B * n, 3 * heads * dh) in `prec` (the qkv GEMM's PP_OUT_OPERAND output);
// out_op: operand (B * n, heads * dh).
int launch_attention_mma(int prec, const void* qkv_op, int batch, int n, int heads, int dh, void* out_op, cudaStream_t st) {
  PP_REQUIRE(attention_mma_supported(n, dh), PP_ERR_UNSUPPORTED,
             "tensor-core attention is built for 192 tokens and head width 32 / 64 (got %d tokens, width %d)", n, dh);
  if (batch == 0) return PP_OK;
  return dh == 32 ? launch_mma_prec<32>(prec, qkv_op, batch, heads, out_op, st)
                  : launch_mma_prec<64>(prec, qkv_op, batch, heads, out_op, st);
}

}  // namespace pp
