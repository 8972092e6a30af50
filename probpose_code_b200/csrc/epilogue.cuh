// GEMM epilogue shared by the tcgen05 kernel and the CUDA-core verification kernel:
//   out = act(acc * scale[n] + shift[n]) (+ residual), written as fp32 rows, as the next
//   GEMM's operand, or as channel-major planes; optional ConvTranspose2d phase scatter.
#pragma once

#include "common.cuh"

namespace pp {

struct EpiParams {
  const float* scale;
  const float* shift;
  const float* residual;
  void* d;
  int m, n;
  int act, out_kind, ldd, plane;
  int up_hin, up_win, up_py, up_px;
  int res_mod;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == PP_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));  // nn.GELU (erf form)
  if (act == PP_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// Logical GEMM row -> output row (identity, or sub-pixel phase scatter of a stride-2 deconv).
__device__ __forceinline__ int64_t epi_out_row(const EpiParams& e, int m) {
  if (e.up_hin == 0) return m;
  const int j = m % e.up_win;
  const int t = m / e.up_win;
  const int i = t % e.up_hin;
  const int b = t / e.up_hin;
  return (int64_t)(b * 2 * e.up_hin + 2 * i + e.up_py) * (2 * e.up_win) + 2 * j + e.up_px;
}

// Finish and store NC consecutive columns [n0, n0 + NC) of logical row m (m < e.m checked by
// the caller).  `sc` / `sh` point at the scale / shift of column n0 (never NULL).
template <int PREC, int NC>
__device__ __forceinline__ void epi_store(const EpiParams& e, int m, int n0, float (&v)[NC], const float* sc,
                                          const float* sh) {
  static_assert(NC % 4 == 0, "column chunk must be a multiple of 4");
  const int64_t orow = epi_out_row(e, m);
  const bool full = (n0 + NC <= e.n);
#pragma unroll
  for (int c = 0; c < NC; ++c) v[c] = apply_act(fmaf(v[c], sc[c], sh[c]), e.act);

  if (e.out_kind == PP_OUT_F32) {
    float* drow = reinterpret_cast<float*>(e.d) + orow * e.ldd + n0;
    const float* rrow = e.residual ? e.residual + (e.res_mod > 0 ? orow % e.res_mod : orow) * e.ldd + n0 : nullptr;
    if (full && (e.ldd & 3) == 0) {
#pragma unroll
      for (int c = 0; c < NC; c += 4) {
        float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        if (rrow) {
          const float4 r = *reinterpret_cast<const float4*>(rrow + c);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(drow + c) = o;
      }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (n0 + c < e.n) drow[c] = v[c] + (rrow ? rrow[c] : 0.f);
    }
  } else if (e.out_kind == PP_OUT_OPERAND) {
    // next GEMM's A operand, logical width ldd (>= n; columns >= n are not touched)
    if (full) {
#pragma unroll
      for (int c = 0; c < NC; c += 4)
        store_operand4<PREC>(e.d, orow, n0 + c, e.ldd, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (n0 + c < e.n) store_operand<PREC>(e.d, orow, n0 + c, e.ldd, v[c]);
    }
  } else {  // PP_OUT_PLANES: (M / plane, N, plane)
    const int64_t img = orow / e.plane, pix = orow % e.plane;
    float* dbase = reinterpret_cast<float*>(e.d) + (img * e.n + n0) * e.plane + pix;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (n0 + c < e.n) dbase[(int64_t)c * e.plane] = v[c];
  }
}

}  // namespace pp
