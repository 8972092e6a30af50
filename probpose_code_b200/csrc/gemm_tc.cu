// tcgen05 tensor-core GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] . W[N,K]^T), fp32 accumulate.
//
// One 128 x BN output tile per CTA, warp-specialised:
//   warp 0      TMA producer  - cp.async.bulk.tensor (128-byte swizzle) of A / W k-blocks into a
//                               STAGES-deep shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer    - one thread issues tcgen05.mma (M=128, N=BN, K=16) per 32-byte
//                               k-slice; accumulators live in TMEM; tcgen05.commit frees the slot
//   warps 2..5  epilogue      - tcgen05.ld 32 lanes x 32 columns -> registers -> scale/shift/
//                               activation/residual -> global (fp32 rows, next operand, or planes)
//
// Precision modes (pp_precision):
//   FP16 / BF16  one MMA per k-slice.
//   FP16X3       operands are [hi | lo*2^11] fp16 pairs; three MMAs per k-slice:
//                acc0 += Ahi.Whi ; acc1 += Ahi.Wlo' + Alo'.Whi ; result = acc0 + acc1 * 2^-11.
//                fp16 products are exact in the fp32 accumulator, so this recovers ~2^-22
//                relative operand precision (fp32-grade) at 1/3 of the fp16 tensor rate.
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace pp {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 x 16-bit = 128 bytes = one swizzle row
constexpr int kGemmThreads = 192;
constexpr int kSmemBudget = 200 * 1024;

template <int BN, int SPLIT>
struct GemmCfg {
  static constexpr int NOPS = (SPLIT == 3) ? 2 : 1;
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE_BYTES = NOPS * (A_BYTES + B_BYTES);
  static constexpr int STAGES_RAW = kSmemBudget / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_COLS = NOPS * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  // ring | full[STAGES] empty[STAGES] tmem_full | tmem ptr | scale[BN] shift[BN] | 1024 B alignment slack
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 2 * BN * 4 + 1024;
  static_assert(STAGES >= 2, "tile too large for a 2-stage ring");
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
};

template <int BN, int SPLIT, bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const int K,
               const EpiParams e) {
  using Cfg = GemmCfg<BN, SPLIT>;
  constexpr int PREC = SPLIT == 3 ? PP_PREC_FP16X3 : (BF16 ? PP_PREC_BF16 : PP_PREC_FP16);
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_ptr + 4);
  float* s_shift = s_scale + BN;

  auto stage_a = [&](int s, int part) { return smem + s * Cfg::STAGE_BYTES + part * Cfg::A_BYTES; };
  auto stage_b = [&](int s, int part) { return smem + s * Cfg::STAGE_BYTES + Cfg::NOPS * Cfg::A_BYTES + part * Cfg::B_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBM;
  const int num_kb = K / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_w);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
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

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        ptx::tma_load_2d(stage_a(s, 0), &tm_a, &full_bar[s], kb * kBK, m0);
        ptx::tma_load_2d(stage_b(s, 0), &tm_w, &full_bar[s], kb * kBK, n0);
        if constexpr (SPLIT == 3) {
          ptx::tma_load_2d(stage_a(s, 1), &tm_a, &full_bar[s], K + kb * kBK, m0);
          ptx::tma_load_2d(stage_b(s, 1), &tm_w, &full_bar[s], K + kb * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BF16, kBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tcgen05_fence_after();
        const uint32_t a_hi = ptx::smem_u32(stage_a(s, 0)), b_hi = ptx::smem_u32(stage_b(s, 0));
#pragma unroll
        for (int kk = 0; kk < kBK / 16; ++kk) {
          const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
          const uint64_t da = ptx::make_sw128_kmajor_desc(a_hi + kk * 32);
          const uint64_t db = ptx::make_sw128_kmajor_desc(b_hi + kk * 32);
          ptx::umma_f16(tmem_base, da, db, idesc, acc);
          if constexpr (SPLIT == 3) {
            const uint64_t da_lo = ptx::make_sw128_kmajor_desc(ptx::smem_u32(stage_a(s, 1)) + kk * 32);
            const uint64_t db_lo = ptx::make_sw128_kmajor_desc(ptx::smem_u32(stage_b(s, 1)) + kk * 32);
            ptx::umma_f16(tmem_base + BN, da, db_lo, idesc, acc);
            ptx::umma_f16(tmem_base + BN, da_lo, db, idesc, 1u);
          }
        }
        ptx::umma_commit(&empty_bar[s]);  // slot reusable once these MMAs have read it
      }
      ptx::umma_commit(tmem_full_bar);  // accumulators complete
    }
  } else {
    // ===== epilogue (128 threads) =====
    const int et = threadIdx.x - 64;
    for (int c = et; c < BN; c += 128) {
      const bool in = n0 + c < e.n;
      s_scale[c] = (in && e.scale) ? e.scale[n0 + c] : 1.f;
      s_shift[c] = (in && e.shift) ? e.shift[n0 + c] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      ptx::tmem_ld_32x32b_x32(lane_base + c0, v);
      if constexpr (SPLIT == 3) {
        float w[32];
        ptx::tmem_ld_32x32b_x32(lane_base + BN + c0, w);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(w[i], kLoScaleInv, v[i]);
      }
      if (row < e.m && n0 + c0 < e.n) epi_store<PREC, 32>(e, row, n0 + c0, v, s_scale + c0, s_shift + c0);
    }
    ptx::tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

// K-major 16-bit operand (rows x row_elems), box = 64 elements x box_rows, 128-byte swizzle, OOB -> 0.
static int make_operand_map(CUtensorMap* out, const void* base, int64_t rows, int64_t row_elems, int box_rows, bool bf16) {
  typedef std::tuple<const void*, int64_t, int64_t, int, bool> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  const Key key(base, rows, row_elems, box_rows, bf16);
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
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PP_REQUIRE(r == CUDA_SUCCESS, PP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld elems=%lld box_rows=%d", (int)r,
             (long long)rows, (long long)row_elems, box_rows);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return PP_OK;
}

template <int BN, int SPLIT, bool BF16>
static int launch_tc(const pp_gemm_args& a, const EpiParams& e, cudaStream_t st) {
  using Cfg = GemmCfg<BN, SPLIT>;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, SPLIT, BF16>;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int64_t row_elems = (SPLIT == 3 ? 2 : 1) * (int64_t)a.k;
  CUtensorMap tma, tmw;
  int rc = make_operand_map(&tma, a.a, a.m, row_elems, kBM, BF16);
  if (rc) return rc;
  rc = make_operand_map(&tmw, a.w, a.n, row_elems, BN, BF16);
  if (rc) return rc;
  dim3 grid((a.n + BN - 1) / BN, (a.m + kBM - 1) / kBM);
  kern<<<grid, kGemmThreads, Cfg::SMEM_BYTES, st>>>(tma, tmw, a.k, e);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

template <int SPLIT, bool BF16>
static int launch_tc_bn(int bn, const pp_gemm_args& a, const EpiParams& e, cudaStream_t st) {
  switch (bn) {
    case 32: return launch_tc<32, SPLIT, BF16>(a, e, st);
    case 64: return launch_tc<64, SPLIT, BF16>(a, e, st);
    case 128: return launch_tc<128, SPLIT, BF16>(a, e, st);
    case 192: return launch_tc<192, SPLIT, BF16>(a, e, st);
    case 256: return launch_tc<256, SPLIT, BF16>(a, e, st);
  }
  set_error("pp_gemm: unsupported tile_n %d", bn);
  return PP_ERR_INVALID;
}

static int pick_tile_n(int n, int prec) {
  if (n <= 32) return 32;
  if (n <= 64) return 64;
  if (prec == PP_PREC_FP16X3) return (n % 128 == 0 || n > 192) ? 128 : 192;
  if (n % 256 == 0) return 256;
  if (n % 192 == 0) return 192;
  return 128;
}

int gemm_tc_launch(const pp_gemm_args& a, const EpiParams& e, int tile_n, cudaStream_t st) {
  PP_REQUIRE(a.k % kBK == 0 && a.k > 0, PP_ERR_INVALID, "pp_gemm: k=%d must be a positive multiple of %d", a.k, kBK);
  const int bn = tile_n > 0 ? tile_n : pick_tile_n(a.n, a.precision);
  switch (a.precision) {
    case PP_PREC_FP16X3: return launch_tc_bn<3, false>(bn, a, e, st);
    case PP_PREC_BF16: return launch_tc_bn<1, true>(bn, a, e, st);
    case PP_PREC_FP16: return launch_tc_bn<1, false>(bn, a, e, st);
  }
  set_error("pp_gemm: precision %d is not a tensor-core mode", a.precision);
  return PP_ERR_INVALID;
}

}  // namespace pp
