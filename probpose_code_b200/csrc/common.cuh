// Shared helpers for the probpose_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <utility>

#include "../../include/probpose_b200.h"

namespace pp {

// ---- error plumbing ------------------------------------------------------------------
void set_error(const char* fmt, ...);  // defined in capi.cu (thread-local buffer)

#define PP_CHECK_CUDA(expr)                                                               \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      pp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PP_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define PP_REQUIRE(cond, status, ...)                                                     \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      pp::set_error(__VA_ARGS__);                                                         \
      return (status);                                                                    \
    }                                                                                     \
  } while (0)

// Launch counter (bench.py's gpu_launches); bumped by every host-side launcher.
extern thread_local int64_t g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }

constexpr int kWarp = 32;

// One-time set-up that is PER DEVICE (function attributes such as MaxDynamicSharedMemorySize belong to a device's
// context; several devices may be driven from one process): first() is true once per CUDA device.
struct PerDeviceOnce {
  std::atomic<uint64_t> done[4] = {};  // devices 0 .. 255
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const uint64_t bit = 1ull << (dev & 63);
    return !(done[(dev >> 6) & 3].fetch_or(bit) & bit);
  }
};
int device_sm_count();  // capi.cu: SM count of the CURRENT device (cached per device)

// ---- programmatic dependent launch -------------------------------------------------------
// Every kernel of the step is launched with programmatic stream serialization: it may start (barrier
// init, TMEM allocation, descriptor prefetch) while the previous kernel drains, and blocks in
// pdl_wait() until that kernel has completed and its writes are visible.  pdl_wait() must precede the
// first access to any global memory another kernel of the stream writes or still reads.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();  // capi.cu: on unless PP_NO_PDL is set in the environment

// ---- L2 residency of the residual stream ----------------------------------------------------
// The fp32 residual stream x (M x D, 38 MB for 64 crops with the flipped pass) is read and rewritten in place by six
// kernels of every ViT layer while ~1 GB of other activations streams through the 126 MB L2 between two of its uses, so
// without help it comes back from HBM every time (proj and the LayerNorms are HBM-bound on exactly these bytes).  The
// engine marks it PERSISTING: every launch between L2Window::set and ::clear carries an access-policy window over x
// (a launch attribute, so it works on any stream including the legacy default one and is captured into graph nodes).
struct L2Window {
  static thread_local cudaAccessPolicyWindow win;  // num_bytes == 0: none
  static void set(const void* base, size_t bytes, float hit_ratio) {
    win.base_ptr = const_cast<void*>(base); win.num_bytes = bytes; win.hitRatio = hit_ratio;
    win.hitProp = cudaAccessPropertyPersisting; win.missProp = cudaAccessPropertyStreaming;
  }
  static void clear() { win.num_bytes = 0; }
  // appends the window to a launch's attribute list; returns the new count
  static int attach(cudaLaunchAttribute* attr, int n) {
    if (win.num_bytes == 0) return n;
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow = win;
    return n + 1;
  }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  na = L2Window::attach(attr, na);
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming 128-bit load that does not pollute L1 (data is touched once).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- GEMM operand formats --------------------------------------------------------------
// An "operand" is the K-major matrix a tensor-core GEMM reads through TMA.
//   FP16X3 : fp16 (rows, 2K): [hi | lo], hi = rn16(64 a), lo = rn16(64 a - hi).  The power-of-two
//            operand scale keeps lo (about 2^-11 |hi|) a NORMAL fp16 number down to |a| = 2^-9
//            (below that its absolute error is <= 2^-31, irrelevant), so hi.hi + hi.lo + lo.hi can
//            be accumulated in ONE fp32 accumulator; products carry 64 * 64 = 2^12, removed
//            exactly in the epilogue.  |a| saturates at 1023.5.
//   BF16   : bf16 (rows, K)
//   FP16   : fp16 (rows, K)
//   FP32   : fp32 (rows, K)   (CUDA-core verification path)
constexpr float kOpScale = 64.0f;             // FP16X3 operand scale
constexpr float kOpScaleInv = 1.0f / 64.0f;
constexpr float kAccScaleInv = 1.0f / 4096.0f;  // (A * 64) . (W * 64) -> A . W

__host__ __device__ inline int operand_elem_bytes(int prec) { return prec == PP_PREC_FP32_SIMT ? 4 : 2; }
__host__ __device__ inline int64_t operand_row_elems(int prec, int64_t k) { return prec == PP_PREC_FP16X3 ? 2 * k : k; }

// ---- operand range guard -----------------------------------------------------------------
// fp16 operands clamp at +-65504 (FP16X3: |a| > 1023.5 after the 64x operand scale).  A clamp is silent wrong
// output, so every producer of an operand raises a sticky device flag when it clamps; pp_operand_overflow() reads
// (and clears) it.  One flag per translation unit - the library is built without relocatable device code - each
// registered with capi.cu at load time.
static __device__ unsigned g_op_overflow;
__device__ __forceinline__ void note_overflow(float v) {
  if (fabsf(v) > 65504.0f) g_op_overflow = 1u;
}
__device__ __forceinline__ void note_overflow4(float a, float b, float c, float d) {
  if (fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), fabsf(d))) > 65504.0f) g_op_overflow = 1u;
}
typedef int (*OverflowReader)(int clear, unsigned* flagged);
void register_overflow_reader(OverflowReader fn);  // capi.cu
namespace {
struct OverflowRegistration {
  OverflowRegistration() {
    register_overflow_reader([](int clear, unsigned* flagged) -> int {
      unsigned v = 0;
      if (cudaMemcpyFromSymbol(&v, g_op_overflow, sizeof(v)) != cudaSuccess) return 1;
      if (v && clear) {
        const unsigned zero = 0;
        if (cudaMemcpyToSymbol(g_op_overflow, &zero, sizeof(zero)) != cudaSuccess) return 1;
      }
      *flagged = v;
      return 0;
    });
  }
};
static OverflowRegistration overflow_registration_;
}  // namespace

// ---- packed fp32 pairs (sm_100: two IEEE fp32 operations per issue slot) ---------------------
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t bc2(float c) { return pk2(c, c); }

__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {  // {lo, hi} -> f16x2, clamped to +-65504
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}


__device__ __forceinline__ __half sat_half(float v) {
  note_overflow(v);
  return __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
}

// Write logical element (row, col) = v of an operand with logical width k.
template <int PREC>
__device__ __forceinline__ void store_operand(void* base, int64_t row, int col, int k, float v) {
  if constexpr (PREC == PP_PREC_FP16X3) {
    __half* p = reinterpret_cast<__half*>(base) + row * (2 * (int64_t)k);
    v *= kOpScale;
    __half hi = sat_half(v);
    p[col] = hi;
    p[k + col] = __float2half_rn(v - __half2float(hi));
  } else if constexpr (PREC == PP_PREC_BF16) {
    reinterpret_cast<__nv_bfloat16*>(base)[row * (int64_t)k + col] = __float2bfloat16_rn(v);
  } else if constexpr (PREC == PP_PREC_FP16) {
    reinterpret_cast<__half*>(base)[row * (int64_t)k + col] = sat_half(v);
  } else {
    reinterpret_cast<float*>(base)[row * (int64_t)k + col] = v;
  }
}

// Vector form: 4 consecutive columns (col % 4 == 0).
template <int PREC>
__device__ __forceinline__ void store_operand4(void* base, int64_t row, int col, int k, float4 v) {
  if constexpr (PREC == PP_PREC_FP16X3) {
    // hi = rn16(sat(64 v)), lo = rn16(64 v - hi), on packed pairs: the same bits as store_operand element by element
    // (cvt.rn.satfinite clamps to +-65504 like sat_half; the residue is exact in fp32)
    __half* p = reinterpret_cast<__half*>(base) + row * (2 * (int64_t)k);
    const uint64_t v01 = mul2(pk2(v.x, v.y), bc2(kOpScale)), v23 = mul2(pk2(v.z, v.w), bc2(kOpScale));
    upk2(v01, v.x, v.y); upk2(v23, v.z, v.w);
    note_overflow4(v.x, v.y, v.z, v.w);
    uint2 hi, lo;
    hi.x = pack_half2_sat(v.x, v.y); hi.y = pack_half2_sat(v.z, v.w);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
    float r0, r1, r2, r3;
    upk2(sub2(v01, pk2(f0.x, f0.y)), r0, r1);
    upk2(sub2(v23, pk2(f1.x, f1.y)), r2, r3);
    const __half2 l0 = __floats2half2_rn(r0, r1), l1 = __floats2half2_rn(r2, r3);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(p + col) = hi;
    *reinterpret_cast<uint2*>(p + k + col) = lo;
  } else if constexpr (PREC == PP_PREC_BF16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + row * (int64_t)k + col) = o;
  } else if constexpr (PREC == PP_PREC_FP16) {
    __half2 a = __halves2half2(sat_half(v.x), sat_half(v.y)), b = __halves2half2(sat_half(v.z), sat_half(v.w));
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(base) + row * (int64_t)k + col) = o;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + row * (int64_t)k + col) = v;
  }
}

// Dispatch a templated-on-precision callable.
#define PP_DISPATCH_PREC(prec, ...)                                   \
  switch (prec) {                                                     \
    case PP_PREC_FP16X3: { constexpr int PREC = PP_PREC_FP16X3; __VA_ARGS__; break; }       \
    case PP_PREC_BF16: { constexpr int PREC = PP_PREC_BF16; __VA_ARGS__; break; }           \
    case PP_PREC_FP16: { constexpr int PREC = PP_PREC_FP16; __VA_ARGS__; break; }           \
    default: { constexpr int PREC = PP_PREC_FP32_SIMT; __VA_ARGS__; break; }                \
  }

}  // namespace pp
