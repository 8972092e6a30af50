// pp_gemm entry point, the CUDA-core fp32 verification GEMM, and operand conversion.
#include "common.cuh"
#include "epilogue.cuh"

namespace pp {

int gemm_tc_launch(const pp_gemm_args& a, const EpiParams& e, int tile_n, cudaStream_t st);  // gemm_tc.cu
int gemm_tc_launch_group(const pp_gemm_args* a, const EpiParams& e0, int count, cudaStream_t st);

// ---- fp32 FFMA GEMM (PP_PREC_FP32_SIMT): 64x64 tile, 256 threads, 4x4 outputs each ----------
// Exists to verify the tensor-core path on the GPU itself (same epilogue, same layouts).
constexpr int kSB = 64, kSK = 16;

struct SimtTaps { int taps, tap_k; int shift[9]; };

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ W, int K,
                                                        const SimtTaps tp, const EpiParams e) {
  __shared__ float sA[kSK][kSB + 4];
  __shared__ float sW[kSK][kSB + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * kSB, n0 = blockIdx.x * kSB;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kSK) {
    for (int i = threadIdx.x; i < kSB * kSK; i += 256) {
      const int r = i / kSK, c = i % kSK;
      const int kg = k0 + c, tap = kg / tp.tap_k;
      const int64_t ar = (int64_t)m0 + r + tp.shift[tap];  // tap operands: shifted source row, zero outside
      sA[c][r] = (m0 + r < e.m && ar >= 0 && ar < e.m) ? A[ar * tp.tap_k + (kg - tap * tp.tap_k)] : 0.f;
      sW[c][r] = (n0 + r < e.n) ? W[(int64_t)(n0 + r) * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; w[i] = sW[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    sc[j] = (n < e.n && e.scale) ? e.scale[n] : 1.f;
    sh[j] = (n < e.n && e.shift) ? e.shift[n] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m < e.m && n0 + tx * 4 < e.n) epi_store<PP_PREC_FP32_SIMT, 4>(e, m, n0 + tx * 4, acc[i], sc, sh);
  }
}

// ---- fp32 -> operand conversion -----------------------------------------------------------------
template <int PREC>
__global__ void operand_from_f32_kernel(const float* __restrict__ src, int64_t rows, int k, int64_t ld, void* dst) {
  const int64_t n4 = rows * (k / 4);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (k / 4);
    const int c = (int)(i % (k / 4)) * 4;
    const float* s = src + r * ld + c;
    store_operand4<PREC>(dst, r, c, k, make_float4(s[0], s[1], s[2], s[3]));
  }
}

}  // namespace pp

extern "C" size_t pp_operand_bytes(int32_t precision, int64_t rows, int64_t k) {
  return (size_t)rows * pp::operand_row_elems(precision, k) * pp::operand_elem_bytes(precision);
}

extern "C" int pp_operand_from_f32(int32_t precision, const float* src, int64_t rows, int64_t k, int64_t ld_src, void* dst,
                                   void* stream) {
  using namespace pp;
  PP_REQUIRE(src && dst, PP_ERR_INVALID, "pp_operand_from_f32: NULL pointer");
  PP_REQUIRE(precision >= 0 && precision <= PP_PREC_FP32_SIMT, PP_ERR_INVALID, "pp_operand_from_f32: bad precision %d", precision);
  PP_REQUIRE(k > 0 && k % 4 == 0 && ld_src >= k, PP_ERR_INVALID, "pp_operand_from_f32: k=%lld ld=%lld", (long long)k, (long long)ld_src);
  if (rows == 0) return PP_OK;
  const int64_t n4 = rows * (k / 4);
  const int grid = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  PP_DISPATCH_PREC(precision, (operand_from_f32_kernel<PREC><<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, (int)k, ld_src, dst)));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

namespace pp {

static EpiParams make_epi(const pp_gemm_args& a) {
  EpiParams e;
  e.scale = a.scale; e.shift = a.shift; e.residual = a.residual; e.d = a.d;
  e.m = a.m; e.n = a.n; e.act = a.act; e.out_kind = a.out_kind; e.ldd = a.ldd; e.plane = a.plane;
  e.up_hin = a.up_hin; e.up_py = a.up_py; e.up_px = a.up_px;
  e.in_pad = a.in_pad; e.out_pad = a.out_pad;
  e.in_h = a.in_pad ? a.in_h : a.up_hin; e.in_w = a.in_pad ? a.in_w : a.up_win;
  if (a.out_pad && !a.in_pad && !a.up_hin) { e.in_h = a.in_h; e.in_w = a.in_w; }
  e.res_mod = a.res_mod;
  return e;
}

int gemm_dispatch(const pp_gemm_args& a, cudaStream_t st);

// `count` GEMMs that share everything but W, the tap shifts and the ConvTranspose2d phase: one persistent launch on the
// tensor-core path (gemm_tc.cu "Grouped launch"), one launch each on the CUDA-core verification path.
int gemm_dispatch_group(const pp_gemm_args* a, int count, cudaStream_t st) {
  if (count <= 0 || a[0].m == 0) return PP_OK;
  if (a[0].precision == PP_PREC_FP32_SIMT || count == 1) {
    for (int g = 0; g < count; ++g) {
      const int rc = gemm_dispatch(a[g], st);
      if (rc != PP_OK) return rc;
    }
    return PP_OK;
  }
  return gemm_tc_launch_group(a, make_epi(a[0]), count, st);
}

int gemm_dispatch(const pp_gemm_args& a, cudaStream_t st) {
  if (a.m == 0) return PP_OK;
  const EpiParams e = make_epi(a);
  if (a.precision == PP_PREC_FP32_SIMT) {
    SimtTaps tp = {};
    tp.taps = a.a_taps > 1 ? a.a_taps : 1;
    PP_REQUIRE(tp.taps <= 9 && a.k % tp.taps == 0, PP_ERR_INVALID, "pp_gemm: %d taps do not divide k=%d", tp.taps, a.k);
    tp.tap_k = a.k / tp.taps;
    for (int t = 0; t < tp.taps; ++t) tp.shift[t] = a.a_taps > 1 ? a.a_tap_shift[t] : 0;
    dim3 grid((a.n + kSB - 1) / kSB, (a.m + kSB - 1) / kSB);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(a.a), reinterpret_cast<const float*>(a.w), a.k, tp, e);
    count_launch();
    PP_CHECK_CUDA(cudaGetLastError());
    return PP_OK;
  }
  return gemm_tc_launch(a, e, a.tile_n, st);
}

}  // namespace pp

extern "C" int pp_gemm(const pp_gemm_args* a, void* stream) {
  using namespace pp;
  PP_REQUIRE(a && a->a && a->w && a->d, PP_ERR_INVALID, "pp_gemm: NULL argument");
  PP_REQUIRE(a->precision >= PP_PREC_FP16X3 && a->precision <= PP_PREC_FP32_SIMT, PP_ERR_INVALID, "pp_gemm: bad precision %d", a->precision);
  PP_REQUIRE(a->m >= 0 && a->n > 0 && a->k > 0, PP_ERR_INVALID, "pp_gemm: bad shape m=%d n=%d k=%d", a->m, a->n, a->k);
  PP_REQUIRE(a->out_kind >= PP_OUT_F32 && a->out_kind <= PP_OUT_PLANES, PP_ERR_INVALID, "pp_gemm: bad out_kind %d", a->out_kind);
  PP_REQUIRE(a->out_kind != PP_OUT_PLANES || a->plane > 0, PP_ERR_INVALID, "pp_gemm: PP_OUT_PLANES needs plane > 0");
  PP_REQUIRE(a->out_kind == PP_OUT_PLANES || a->ldd >= a->n, PP_ERR_INVALID, "pp_gemm: ldd=%d < n=%d", a->ldd, a->n);
  PP_REQUIRE(a->out_kind != PP_OUT_OPERAND || (a->ldd % 4 == 0 && a->residual == nullptr), PP_ERR_INVALID,
             "pp_gemm: operand output needs ldd %% 4 == 0 and no residual");
  PP_REQUIRE(a->res_mod >= 0, PP_ERR_INVALID, "pp_gemm: res_mod=%d", a->res_mod);
  PP_REQUIRE(a->res_mod == 0 || (!a->up_hin && !a->in_pad && !a->out_pad), PP_ERR_INVALID,
             "pp_gemm: res_mod needs the identity row mapping");
  PP_REQUIRE(a->a_taps >= 0 && a->a_taps <= 9, PP_ERR_INVALID, "pp_gemm: a_taps=%d outside [0, 9]", a->a_taps);
  PP_REQUIRE(a->in_pad >= 0 && a->in_pad <= 2 && a->out_pad >= 0 && a->out_pad <= 2, PP_ERR_INVALID,
             "pp_gemm: in_pad=%d / out_pad=%d outside {0, 1, 2}", a->in_pad, a->out_pad);
  PP_REQUIRE((!a->in_pad && !(a->out_pad && !a->up_hin)) || (a->in_h > 0 && a->in_w > 0), PP_ERR_INVALID,
             "pp_gemm: in_pad / out_pad need in_h, in_w > 0");
  PP_REQUIRE(!a->up_hin || a->up_win > 0, PP_ERR_INVALID, "pp_gemm: up_hin without up_win");
  PP_REQUIRE(!(a->up_hin && a->in_pad) || (a->up_hin == a->in_h && a->up_win == a->in_w), PP_ERR_INVALID,
             "pp_gemm: up_hin/up_win must equal in_h/in_w when in_pad is set");
  return gemm_dispatch(*a, (cudaStream_t)stream);
}
