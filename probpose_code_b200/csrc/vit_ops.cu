// Non-GEMM kernels of the ViT backbone: patch extraction (+ preprocessing, + mirrored pass),
// LayerNorm -> GEMM operand, global attention, layout changes.
//
// Reference semantics: PoseDataPreprocessor (data_preprocessor.py:79-104), mmpretrain 1.2.0
// VisionTransformer / MultiheadAttention (config :56-67), mmcv PatchEmbed (in-tree twin
// mmpose/models/utils/transformer.py:153-245), flipped pass topdown.py:109-112.
#include "engine_ops.cuh"
#include "ln_row.cuh"

#include <math.h>

namespace pp {

// ---- patch extraction --------------------------------------------------------------------
template <int PREC>
__global__ void __launch_bounds__(256) patchify_kernel(const PatchifyParams p, void* a_op) {
  pdl_launch_dependents();
  pdl_wait();
  const int P = p.patch, PP = P * P, K = 3 * PP, K4 = K / 4;
  const int tokens = p.gh * p.gw;
  const int64_t total = (int64_t)p.passes * p.batch * tokens * K4;
  const size_t plane = (size_t)p.img_h * p.img_w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K4) * 4;
    const int64_t m = i / K4;
    const int tok = (int)(m % tokens);
    const int bb = (int)(m / tokens);
    const int pass = bb / p.batch, b = bb % p.batch;
    const int c = k / PP, ky = (k % PP) / P, kx = k % P;
    const int y = (tok / p.gw) * P - p.pad + ky;
    const int x0 = (tok % p.gw) * P - p.pad + kx;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < p.img_h) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int x = x0 + e;                      // column in the (possibly mirrored) image
        if (x < 0 || x >= p.img_w) continue; // zero border of the NORMALISED image
        if (pass) x = p.img_w - 1 - x;
        if (p.u8_bgr) {
          const float raw = (float)p.u8_bgr[((size_t)b * 3 + (2 - c)) * plane + (size_t)y * p.img_w + x];
          v[e] = (raw - p.mean[c]) * p.inv_std[c];
        } else {
          v[e] = p.x_f32[((size_t)b * 3 + c) * plane + (size_t)y * p.img_w + x];
        }
      }
    }
    store_operand4<PREC>(a_op, m, k, K, make_float4(v[0], v[1], v[2], v[3]));
  }
}

// Patch size 16 (every shipped config): one thread per (token, channel, patch row) = 16 pixels of one image row ->
// 16 consecutive operand columns (two 16-byte stores per plane, a warp writes 1 KB runs).  Index arithmetic by
// constants except one division pair per thread; 3 x fewer instructions than the generic kernel above.
template <int PREC>
__global__ void __launch_bounds__(256) patchify16_kernel(const PatchifyParams p, void* a_op) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int P = 16, K = 3 * P * P;
  const int tokens = p.gh * p.gw;
  const int64_t total = (int64_t)p.passes * p.batch * tokens * (3 * P);
  const size_t plane = (size_t)p.img_h * p.img_w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % (3 * P));  // c * 16 + ky
    const int64_t m = i / (3 * P);
    const int tok = (int)(m % tokens), bb = (int)(m / tokens);
    const int pass = bb >= p.batch ? 1 : 0, b = bb - pass * p.batch;
    const int c = r >> 4, ky = r & 15;
    const int ty = tok / p.gw, tx = tok - ty * p.gw;
    const int y = ty * P - p.pad + ky, x0 = tx * P - p.pad;
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
    if (y >= 0 && y < p.img_h) {
      if (p.u8_bgr) {
        const uint8_t* row = p.u8_bgr + ((size_t)b * 3 + (2 - c)) * plane + (size_t)y * p.img_w;
        const float mean = p.mean[c], inv = p.inv_std[c];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int x = x0 + e;  // column in the (possibly mirrored) image; outside: zero border of the NORMALISED image
          if (x >= 0 && x < p.img_w) v[e] = ((float)__ldg(row + (pass ? p.img_w - 1 - x : x)) - mean) * inv;
        }
      } else {
        const float* row = p.x_f32 + ((size_t)b * 3 + c) * plane + (size_t)y * p.img_w;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int x = x0 + e;
          if (x >= 0 && x < p.img_w) v[e] = __ldg(row + (pass ? p.img_w - 1 - x : x));
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 16; e += 4) store_operand4<PREC>(a_op, m, r * P + e, K, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
  }
}

// uint8 crops, patch size 16, width a multiple of 4 and at most kStageW: one CTA per (pass, image, token row).  The 16
// image rows x 3 channels the token row needs are staged in shared memory with coalesced 32-bit loads (the
// per-thread byte gathers of the kernels above cost one L1 wavefront per lane); row pitch 4 * (W / 4 + 1) bytes keeps
// the (channel, patch row) items of a warp on distinct banks.  Every thread produces half a patch row = one 16-byte
// store per operand plane, consecutive lanes consecutive pieces: a warp store is one 512-byte run (8-byte stores at a
// 32-byte stride - one partly filled sector per lane - made the first version of this kernel store-bound).
constexpr int kStageW = 256;

// Eight consecutive columns (col % 8 == 0) of one operand row: one 16-byte store per plane (same arithmetic as
// store_operand4, common.cuh).
template <int PREC>
__device__ __forceinline__ void store_operand8(void* base, int64_t row, int col, int k, const float (&v)[8]) {
  if constexpr (PREC == PP_PREC_FP16X3) {
    __half* p = reinterpret_cast<__half*>(base) + row * (2 * (int64_t)k);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = v[2 * i] * kOpScale, b = v[2 * i + 1] * kOpScale;
      const __half h0 = sat_half(a), h1 = sat_half(b);
      const __half2 h = __halves2half2(h0, h1);
      const __half2 l = __floats2half2_rn(a - __half2float(h0), b - __half2float(h1));
      hi[i] = *reinterpret_cast<const uint32_t*>(&h);
      lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(p + col) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p + k + col) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  } else {
    store_operand4<PREC>(base, row, col, k, make_float4(v[0], v[1], v[2], v[3]));
    store_operand4<PREC>(base, row, col + 4, k, make_float4(v[4], v[5], v[6], v[7]));
  }
}
template <int PREC>
__global__ void __launch_bounds__(256) patchify16_u8_kernel(const PatchifyParams p, void* a_op) {
  constexpr int P = 16, K = 3 * P * P;
  __shared__ uint32_t stage[3 * P][kStageW / 4 + 1];
  pdl_launch_dependents();
  pdl_wait();
  const int ty = blockIdx.x % p.gh, bb = blockIdx.x / p.gh;
  const int pass = bb >= p.batch ? 1 : 0, b = bb - pass * p.batch;
  const int w4 = p.img_w >> 2;
  const size_t plane = (size_t)p.img_h * p.img_w;
  const int y_first = ty * P - p.pad;
  for (int i = threadIdx.x; i < 3 * P * w4; i += blockDim.x) {
    const int r = i / w4, xw = i - r * w4;  // r = c * 16 + ky
    const int y = y_first + (r & 15);
    if (y >= 0 && y < p.img_h)
      stage[r][xw] = __ldg(reinterpret_cast<const uint32_t*>(p.u8_bgr + ((size_t)b * 3 + (2 - (r >> 4))) * plane + (size_t)y * p.img_w) + xw);
  }
  __syncthreads();
  const int64_t m0 = ((int64_t)bb * p.gh + ty) * p.gw;
  // item = (token, channel * 16 + patch row, half row): consecutive lanes write consecutive 16-byte pieces
  for (int item = threadIdx.x; item < p.gw * 3 * P * 2; item += blockDim.x) {
    const int half = item & 1, tr = item >> 1;
    const int tx = tr / (3 * P), r = tr - tx * (3 * P);
    const int c = r >> 4, y = y_first + (r & 15);
    const bool row_ok = y >= 0 && y < p.img_h;
    const uint8_t* row = reinterpret_cast<const uint8_t*>(stage[r]);
    const float mean = p.mean[c], inv = p.inv_std[c];
    const int x0 = tx * P - p.pad + half * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int x = x0 + e;  // outside the image: zero border of the NORMALISED image
      v[e] = (row_ok && x >= 0 && x < p.img_w) ? ((float)row[pass ? p.img_w - 1 - x : x] - mean) * inv : 0.f;
    }
    store_operand8<PREC>(a_op, m0 + tx, r * P + half * 8, K, v);
  }
}

int launch_patchify(int prec, const PatchifyParams& p, void* a_op, cudaStream_t st) {
  PP_REQUIRE((p.u8_bgr != nullptr) != (p.x_f32 != nullptr), PP_ERR_INVALID,
             "exactly one of crops_u8_bgr and x_f32 must be given");
  PP_REQUIRE(p.patch % 4 == 0, PP_ERR_UNSUPPORTED, "patch size %d not a multiple of 4", p.patch);
  const int64_t total = (int64_t)p.passes * p.batch * p.gh * p.gw * (3 * p.patch * p.patch / 4);
  if (total == 0) return PP_OK;
  if (p.patch == 16 && p.u8_bgr != nullptr && p.img_w % 4 == 0 && p.img_w <= kStageW &&
      (reinterpret_cast<uintptr_t>(p.u8_bgr) & 3) == 0 && ((size_t)p.img_h * p.img_w) % 4 == 0) {
    const int ctas = p.passes * p.batch * p.gh;
    { cudaError_t lerr = cudaSuccess; PP_DISPATCH_PREC(prec, (lerr = launch_pdl(patchify16_u8_kernel<PREC>, dim3(ctas), dim3(256), 0, st, p, a_op))); PP_CHECK_CUDA(lerr); }
    count_launch();
    PP_CHECK_CUDA(cudaGetLastError());
    return PP_OK;
  }
  if (p.patch == 16) {
    const int64_t rows = (int64_t)p.passes * p.batch * p.gh * p.gw * 48;
    const int grid16 = (int)((rows + 255) / 256 < 148 * 32 ? (rows + 255) / 256 : 148 * 32);
    { cudaError_t lerr = cudaSuccess; PP_DISPATCH_PREC(prec, (lerr = launch_pdl(patchify16_kernel<PREC>, dim3(grid16), dim3(256), 0, st, p, a_op))); PP_CHECK_CUDA(lerr); }
    count_launch();
    PP_CHECK_CUDA(cudaGetLastError());
    return PP_OK;
  }
  const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  { cudaError_t lerr = cudaSuccess; PP_DISPATCH_PREC(prec, (lerr = launch_pdl(patchify_kernel<PREC>, dim3(grid), dim3(256), 0, st, p, a_op))); PP_CHECK_CUDA(lerr); }
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

// ---- LayerNorm ----------------------------------------------------------------------------
// One warp per row (ln_row.cuh: the same routine finishes rows inside the tcgen05 GEMM epilogue when the LayerNorm
// is fused into the GEMM that produces its input).
template <int PREC, int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int64_t rows,
                                                        void* out_op, float* out_f32, int pad_gh, int pad_gw) {
  constexpr int D = NV * 128;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  LnParams p;
  p.gamma = gamma; p.beta = beta; p.eps = eps; p.out_op = out_op; p.out_f32 = out_f32; p.pad_gh = pad_gh; p.pad_gw = pad_gw;
  ln_finish_row<PREC, NV>(v, p, row, lane);
}

int launch_layernorm(int prec, const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d,
                     void* out_op, float* out_f32, cudaStream_t st, int pad_gh, int pad_gw) {
  PP_REQUIRE(d == 384 || d == 768, PP_ERR_UNSUPPORTED, "LayerNorm width %d not built (384, 768)", d);
  if (rows == 0) return PP_OK;
  const int grid = (int)((rows + 7) / 8);
  if (d == 384) {
    { cudaError_t lerr = cudaSuccess; PP_DISPATCH_PREC(prec, (lerr = launch_pdl(layernorm_kernel<PREC, 3>, dim3(grid), dim3(256), 0, st, x, gamma, beta, eps, rows, out_op, out_f32, pad_gh, pad_gw))); PP_CHECK_CUDA(lerr); }
  } else {
    { cudaError_t lerr = cudaSuccess; PP_DISPATCH_PREC(prec, (lerr = launch_pdl(layernorm_kernel<PREC, 6>, dim3(grid), dim3(256), 0, st, x, gamma, beta, eps, rows, out_op, out_f32, pad_gh, pad_gw))); PP_CHECK_CUDA(lerr); }
  }
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

// ---- attention (fp32 CUDA-core version) ---------------------------------------------------
// One CTA per (image, head); K and V of the head sit in shared memory, each thread owns one
// query row and walks all keys with a running-max softmax (exact up to fp32 rounding).
template <int PREC, int DH>
__global__ void __launch_bounds__(256) attention_simt_kernel(const float* __restrict__ qkv, int n, int heads,
                                                             void* out_op) {
  extern __shared__ __align__(16) float att_smem[];
  float* sK = att_smem;
  float* sV = att_smem + (size_t)n * DH;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int D = heads * DH;
  const float* base = qkv + (size_t)b * n * 3 * D + h * DH;
  for (int i = threadIdx.x; i < n * (DH / 4); i += blockDim.x) {
    const int r = i / (DH / 4), c4 = i % (DH / 4);
    reinterpret_cast<float4*>(sK)[i] = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * D + D + c4 * 4);
    reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * D + 2 * D + c4 * 4);
  }
  __syncthreads();
  const float scale = rsqrtf((float)DH);
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    float q[DH], o[DH];
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      const float4 v = *reinterpret_cast<const float4*>(base + (size_t)t * 3 * D + d);
      q[d] = v.x * scale; q[d + 1] = v.y * scale; q[d + 2] = v.z * scale; q[d + 3] = v.w * scale;
      o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
    }
    float mx = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < n; j0 += 4) {
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        float acc = 0.f;
        if (j < n) {
          const float4* kr = reinterpret_cast<const float4*>(sK + (size_t)j * DH);
#pragma unroll
          for (int d4 = 0; d4 < DH / 4; ++d4) {
            const float4 kk = kr[d4];
            acc = fmaf(q[4 * d4], kk.x, acc); acc = fmaf(q[4 * d4 + 1], kk.y, acc);
            acc = fmaf(q[4 * d4 + 2], kk.z, acc); acc = fmaf(q[4 * d4 + 3], kk.w, acc);
          }
        } else {
          acc = -INFINITY;
        }
        s[u] = acc;
      }
      const float cm = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
      if (cm > mx) {
        const float r = expf(mx - cm);  // exp(-inf) = 0 on the first chunk
        l *= r;
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] *= r;
        mx = cm;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j >= n) continue;
        const float pj = expf(s[u] - mx);
        l += pj;
        const float4* vr = reinterpret_cast<const float4*>(sV + (size_t)j * DH);
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 vv = vr[d4];
          o[4 * d4] = fmaf(pj, vv.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(pj, vv.y, o[4 * d4 + 1]);
          o[4 * d4 + 2] = fmaf(pj, vv.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(pj, vv.w, o[4 * d4 + 3]);
        }
      }
    }
    const float inv = 1.0f / l;
    const int64_t row = (int64_t)b * n + t;
#pragma unroll
    for (int d = 0; d < DH; d += 4)
      store_operand4<PREC>(out_op, row, h * DH + d, D, make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv));
  }
}

template <int PREC, int DH>
static int launch_attention_t(const float* qkv, int batch, int n, int heads, void* out_op, cudaStream_t st) {
  auto kern = attention_simt_kernel<PREC, DH>;
  const size_t smem = (size_t)2 * n * DH * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int threads = n >= 256 ? 256 : ((n + 31) / 32) * 32;
  kern<<<batch * heads, threads, smem, st>>>(qkv, n, heads, out_op);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

int launch_attention(int prec, const float* qkv, int batch, int n, int heads, int dh, void* out_op, cudaStream_t st) {
  PP_REQUIRE(dh == 32 || dh == 64, PP_ERR_UNSUPPORTED, "attention head width %d not built (32, 64)", dh);
  PP_REQUIRE((size_t)2 * n * dh * 4 <= 200 * 1024, PP_ERR_UNSUPPORTED, "attention: %d tokens do not fit shared memory", n);
  if (batch == 0) return PP_OK;
  int rc = PP_OK;
  if (dh == 32) { PP_DISPATCH_PREC(prec, (rc = launch_attention_t<PREC, 32>(qkv, batch, n, heads, out_op, st))); }
  else          { PP_DISPATCH_PREC(prec, (rc = launch_attention_t<PREC, 64>(qkv, batch, n, heads, out_op, st))); }
  return rc;
}

// ---- layout changes -----------------------------------------------------------------------
// (B * hw, c) rows -> (B, c, hw): 32 x 32 shared-memory transpose tiles.
__global__ void __launch_bounds__(256) rows_to_nchw_kernel(const float* __restrict__ rows, int hw, int c, float* nchw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (p0 + r < hw && c0 + tx < c) tile[r][tx] = rows[((size_t)b * hw + p0 + r) * c + c0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (c0 + r < c && p0 + tx < hw) nchw[((size_t)b * c + c0 + r) * hw + p0 + tx] = tile[tx][r];
}

int launch_rows_to_nchw(const float* rows, int batch, int hw, int c, float* nchw, cudaStream_t st) {
  if (batch == 0) return PP_OK;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, batch);
  rows_to_nchw_kernel<<<grid, 256, 0, st>>>(rows, hw, c, nchw);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

template <int PREC>
__global__ void __launch_bounds__(256) nchw_to_operand_kernel(const float* __restrict__ nchw, int hw, int c, void* rows_op,
                                                              int pad_gh, int pad_gw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (c0 + r < c && p0 + tx < hw) tile[r][tx] = nchw[((size_t)b * c + c0 + r) * hw + p0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (p0 + r < hw && c0 + tx < c) {
      const int64_t row = (int64_t)b * hw + p0 + r;
      store_operand<PREC>(rows_op, pad_gw > 0 ? padded_row(row, pad_gh, pad_gw) : row, c0 + tx, c, tile[tx][r]);
    }
}

int launch_nchw_to_operand(int prec, const float* nchw, int batch, int hw, int c, void* rows_op, cudaStream_t st, int pad_gh,
                           int pad_gw) {
  if (batch == 0) return PP_OK;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, batch);
  PP_DISPATCH_PREC(prec, (nchw_to_operand_kernel<PREC><<<grid, 256, 0, st>>>(nchw, hw, c, rows_op, pad_gh, pad_gw)));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

}  // namespace pp
