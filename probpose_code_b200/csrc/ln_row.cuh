// LayerNorm of one row by one warp, shared by layernorm_kernel (vit_ops.cu) and the fused row finisher of the
// tcgen05 GEMM epilogue (gemm_tc.cu) so that both produce the same bits: the row lives in registers (NV float4 per
// lane, d = 128 * NV), two-pass mean / variance in fp32 like ATen's CPU and CUDA kernels.
#pragma once

#include "common.cuh"

namespace pp {

// pad_gw > 0: the operand output is a shared-border (pad_gh + 1) x (pad_gw + 1) map per image (epilogue.cuh pad_geom
// mode 2: the tap operand of the head's implicit-GEMM convolutions); token (y, x) lands at (y + 1, x).
__device__ __forceinline__ int64_t padded_row(int64_t row, int gh, int gw) {
  const int tokens = gh * gw;
  const int64_t b = row / tokens;
  const int t = (int)(row % tokens);
  return (b * (gh + 1) + t / gw + 1) * (gw + 1) + t % gw;
}

struct LnParams {
  const float* gamma;
  const float* beta;
  float eps;
  void* out_op;     // next GEMM's operand (rows, d), or NULL
  float* out_f32;   // fp32 (rows, d), or NULL
  int pad_gh, pad_gw;
};

// v: the row (lane holds float4 number lane + 32 i of it).  Packed fp32 pairs throughout (two IEEE operations per
// issue slot): inside the GEMM epilogue this runs on warps that have their own tiles to store.
template <int PREC, int NV>
__device__ __forceinline__ void ln_finish_row(const float4 (&v)[NV], const LnParams& p, int64_t row, int lane) {
  constexpr int D = NV * 128;
  uint64_t acc = bc2(0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) acc = add2(acc, add2(pk2(v[i].x, v[i].y), pk2(v[i].z, v[i].w)));
  float s0, s1;
  upk2(acc, s0, s1);
  const uint64_t mean2 = bc2(warp_sum(s0 + s1) * (1.0f / D));
  uint64_t d[NV][2];  // v - mean
  acc = bc2(0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    d[i][0] = sub2(pk2(v[i].x, v[i].y), mean2);
    d[i][1] = sub2(pk2(v[i].z, v[i].w), mean2);
    acc = fma2(d[i][0], d[i][0], acc);
    acc = fma2(d[i][1], d[i][1], acc);
  }
  upk2(acc, s0, s1);
  const uint64_t rstd2 = bc2(rsqrtf(warp_sum(s0 + s1) * (1.0f / D) + p.eps));
  const int64_t op_row = p.pad_gw > 0 ? padded_row(row, p.pad_gh, p.pad_gw) : row;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (lane + 32 * i) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + col));
    float4 o;  // ((v - mean) * rstd) * g + b
    upk2(fma2(mul2(d[i][0], rstd2), pk2(g.x, g.y), pk2(b.x, b.y)), o.x, o.y);
    upk2(fma2(mul2(d[i][1], rstd2), pk2(g.z, g.w), pk2(b.z, b.w)), o.z, o.w);
    if (p.out_op) store_operand4<PREC>(p.out_op, op_row, col, D, o);
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + row * D + col) = o;
  }
}

}  // namespace pp
