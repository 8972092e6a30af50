// Standalone probe (not part of the library): what does a read-only streaming pass over N MB reach on
// this GPU, for a few access structures?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ld_na(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float max4(float4 q) { return fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)); }

// warp per 12 KB map, CH loads in flight per lane and chunk; NA: bypass L1
template <int CH, bool NA>
__global__ void __launch_bounds__(256) k_warp_map(const float* __restrict__ src, float* out, int count) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * 8 + warp;
  if (item >= count) return;
  const float4* g4 = reinterpret_cast<const float4*>(src + (size_t)item * 3072) + lane;
  float mx = -1e30f;
#pragma unroll
  for (int c = 0; c < 24; c += CH) {
    float4 q[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) q[i] = NA ? ld_na(g4 + 32 * (c + i)) : __ldg(g4 + 32 * (c + i));
#pragma unroll
    for (int i = 0; i < CH; ++i) mx = fmaxf(mx, max4(q[i]));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) out[item] = mx;
}

// persistent grid-stride
template <int U>
__global__ void __launch_bounds__(512) k_grid_stride(const float* __restrict__ src, float* out, size_t n4) {
  const float4* g4 = reinterpret_cast<const float4*>(src);
  float mx = -1e30f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = ld_na(g4 + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) mx = fmaxf(mx, max4(q[u]));
  }
  for (; i < n4; i += stride) mx = fmaxf(mx, max4(ld_na(g4 + i)));
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) out[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = mx;
}

// TMA bulk: persistent, W warps per CTA, each with a 12 KB tile; maps taken round-robin
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_bulk(const float* __restrict__ src, float* out, int count) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* tile = reinterpret_cast<float*>(smem) + (threadIdx.x >> 5) * 3072;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + WARPS * 12288);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile);
  if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  uint32_t phase = 0;
  for (int item = blockIdx.x * WARPS + warp; item < count; item += gridDim.x * WARPS) {
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(12288) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src + (size_t)item * 3072), "r"(12288), "r"(bar) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    phase ^= 1;
    const float4* t4 = reinterpret_cast<const float4*>(tile);
    float mx = -1e30f;
#pragma unroll
    for (int j = 0; j < 24; ++j) mx = fmaxf(mx, max4(t4[lane + 32 * j]));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) out[item] = mx;
    __syncwarp();
  }
}

__global__ void k_read_flush(const float4* p, size_t n4, float* out) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { float4 q = p[i]; acc += q.x + q.y + q.z + q.w; }
  if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
  const int sizes[] = {64 * 17, 256 * 17, 512 * 17, 1024 * 17};
  float *src, *out, *flush;
  const size_t max_maps = 1024 * 17;
  CK(cudaMalloc(&src, max_maps * 12288));
  CK(cudaMalloc(&out, max_maps * 4 + (1 << 20)));
  const size_t flush_bytes = 512ull << 20;
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMemset(src, 0, max_maps * 12288));
  CK(cudaMemset(flush, 0, flush_bytes));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaFuncSetAttribute(k_bulk<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 12288 + 256));
  CK(cudaFuncSetAttribute(k_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12288 + 256));
  for (int count : sizes) {
    const double mb = count * 12288 / 1e6;
    auto run = [&](const char* name, auto launch) {
      std::vector<float> t;
      for (int r = 0; r < 25; ++r) {
        k_read_flush<<<1184, 256>>>(reinterpret_cast<const float4*>(flush), flush_bytes / 16, out);  // clean (read-only) L2 eviction
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        t.push_back(ms * 1e3f);
      }
      CK(cudaGetLastError());
      std::sort(t.begin(), t.end());
      printf("%-34s maps %6d  %7.1f MB  best %6.2f us  median %6.2f us  -> %6.0f GB/s (median)\n", name, count, mb, t[0], t[t.size() / 2], mb * 1e6 / t[t.size() / 2] / 1e3);
    };
    const int grid8 = (count + 7) / 8;
    run("empty-ish (1 CTA) launch", [&] { k_warp_map<6, false><<<1, 256>>>(src, out, 8); });
    run("warp/map ldg ch6", [&] { k_warp_map<6, false><<<grid8, 256>>>(src, out, count); });
    run("warp/map ldg ch12", [&] { k_warp_map<12, false><<<grid8, 256>>>(src, out, count); });
    run("warp/map ldg ch24", [&] { k_warp_map<24, false><<<grid8, 256>>>(src, out, count); });
    run("warp/map no_allocate ch6", [&] { k_warp_map<6, true><<<grid8, 256>>>(src, out, count); });
    run("warp/map no_allocate ch12", [&] { k_warp_map<12, true><<<grid8, 256>>>(src, out, count); });
    run("warp/map no_allocate ch24", [&] { k_warp_map<24, true><<<grid8, 256>>>(src, out, count); });
    run("grid-stride 148x512 U4", [&] { k_grid_stride<4><<<148, 512>>>(src, out, (size_t)count * 768); });
    run("grid-stride 296x512 U4", [&] { k_grid_stride<4><<<296, 512>>>(src, out, (size_t)count * 768); });
    run("grid-stride 592x512 U8", [&] { k_grid_stride<8><<<592, 512>>>(src, out, (size_t)count * 768); });
    run("bulk 16 tiles/SM persistent", [&] { k_bulk<16><<<148, 512, 16 * 12288 + 256>>>(src, out, count); });
    run("bulk 2x8 tiles/SM persistent", [&] { k_bulk<8><<<296, 256, 8 * 12288 + 256>>>(src, out, count); });
  }
  return 0;
}
