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
  int up_hin, up_py, up_px;        // up_hin != 0: stride-2 ConvTranspose2d phase scatter
  int in_h, in_w, in_pad, out_pad;  // geometry of the map the GEMM rows enumerate (see map_out_row)
  int res_mod;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == PP_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));  // nn.GELU (erf form)
  if (act == PP_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// Border geometry of a padded map.  mode 1: a one-pixel zero border on every side, (h + 2) x (w + 2) per image,
// pixel (i, j) at (i + 1, j + 1).  mode 2 ("shared border"): (h + 1) x (w + 1) per image - row 0 of every image
// block is zero (top border, and bottom border of the image before; rows past the last block read as zero through
// the TMA out-of-bounds fill) and column w is zero (right border, and left border of the next row: (r, -1) is the
// same element as (r - 1, w)); pixel (i, j) at (i + 1, j).  15 % extra rows instead of 31 % for a 16 x 12 map.
struct PadGeom {
  int xo, yo, ex;  // interior origin, extra rows / columns per image
};
__host__ __device__ __forceinline__ PadGeom pad_geom(int mode) {
  PadGeom g;
  g.xo = mode == 1 ? 1 : 0;
  g.yo = mode != 0 ? 1 : 0;
  g.ex = mode == 1 ? 2 : (mode == 2 ? 1 : 0);
  return g;
}

// Logical GEMM row -> output row and validity.
//   in_pad : the A rows enumerate a zero-padded map (pad_geom); border rows produce nothing
//   up     : stride-2 ConvTranspose2d phase scatter (i, j) -> (2 i + py, 2 j + px)
//   out_pad: the output map carries a zero border of its own (it feeds the next tap GEMM)
__device__ __forceinline__ int64_t map_out_row(const EpiParams& e, int m, bool& valid) {
  valid = m < e.m;
  if (e.up_hin == 0 && e.in_pad == 0 && e.out_pad == 0) return m;
  int b, i, j;
  if (e.in_pad) {
    const PadGeom ig = pad_geom(e.in_pad);
    const int wp = e.in_w + ig.ex, hp = e.in_h + ig.ex;
    const int jp = m % wp, t = m / wp;
    const int ip = t % hp;
    b = t / hp;
    valid = valid && jp >= ig.xo && jp < e.in_w + ig.xo && ip >= ig.yo && ip < e.in_h + ig.yo;
    i = ip - ig.yo; j = jp - ig.xo;
  } else {
    j = m % e.in_w;
    const int t = m / e.in_w;
    i = t % e.in_h;
    b = t / e.in_h;
  }
  const PadGeom og = pad_geom(e.out_pad);
  if (e.up_hin)
    return (int64_t)(b * (2 * e.in_h + og.ex) + 2 * i + e.up_py + og.yo) * (2 * e.in_w + og.ex) + 2 * j + e.up_px + og.xo;
  return (int64_t)(b * (e.in_h + og.ex) + i + og.yo) * (e.in_w + og.ex) + j + og.xo;
}

// Finish and store NC consecutive columns [n0, n0 + NC) of logical row m (m < e.m checked by
// the caller).  `sc` / `sh` point at the scale / shift of column n0 (never NULL).
template <int PREC, int NC>
__device__ __forceinline__ void epi_store(const EpiParams& e, int m, int n0, float (&v)[NC], const float* sc,
                                          const float* sh) {
  static_assert(NC % 4 == 0, "column chunk must be a multiple of 4");
  bool row_ok;
  const int64_t orow = map_out_row(e, m, row_ok);
  if (!row_ok) return;
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
