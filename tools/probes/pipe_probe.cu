// Issue-rate probe for the conversion / special-function instructions of the GEMM and attention epilogues
// (sm_100a): warp instructions per cycle per SM for ex2, rcp, f32x2 -> f16x2 pack, f16 -> f32 unpack, fma, lop3, prmt.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/pipe_probe tools/probes/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(1024) probe(float* out, long long* cyc, float seed) {
  float v[CH];
  uint32_t u[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { v[i] = seed + threadIdx.x * 1e-3f + i; u[i] = __float_as_uint(v[i]); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("{.reg .b32 t; cvt.rn.f16x2.f32 t, %0, %0; and.b32 %0, t, 0x3fff3fff;}" : "+f"(v[i]));  // + 1 alu op
      if (OP == 3) asm volatile("{.reg .f16 lo, hi; .reg .f32 f; mov.b32 {lo, hi}, %0; cvt.f32.f16 f, lo; mov.b32 %0, f;}" : "+r"(u[i]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
      if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
      if (OP == 7) asm volatile("add.f32 %0, %0, %0;" : "+f"(v[i]));
      if (OP == 8) asm volatile("{.reg .b32 t; cvt.rn.bf16x2.f32 t, %0, %0; and.b32 %0, t, 0x3fff3fff;}" : "+f"(v[i]));
      if (OP == 9) asm volatile("{.reg .b32 t; and.b32 t, %0, 0x3fff3fff; and.b32 %0, t, 0x3fffffff;}" : "+r"(u[i]));  // 2 alu ops (baseline for the +1 above)
      if (OP == 10) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 1) % CH]));
      if (OP == 11) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
      if (OP == 12) asm volatile("shf.r.wrap.b32 %0, %0, %1, 13;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
      if (OP == 13) asm volatile("{.reg .f16x2 h; mov.b32 h, %1; fma.rn.f16x2 h, h, h, h; mov.b32 %0, h;}" : "=r"(u[i]) : "r"(u[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) acc += v[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, float* out, long long* cyc) {
  const int blocks = 148, threads = 1024;
  probe<OP><<<blocks, threads>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  probe<OP><<<blocks, threads>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; ++i) mean += h[i];
  mean /= blocks;
  const double warp_instr = (double)ITERS * CH * (threads / 32);
  printf("{\"op\": \"%s\", \"cycles\": %.0f, \"warp_instr_per_clk_per_sm\": %.3f, \"lanes_per_clk_per_sm\": %.1f}\n", name, mean,
         warp_instr / mean, warp_instr * 32 / mean);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<0>("ex2.approx", out, cyc);
  run<1>("rcp.approx", out, cyc);
  run<2>("cvt.rn.f16x2.f32 (pack) + and", out, cyc);
  run<9>("and + and", out, cyc);
  run<8>("cvt.rn.bf16x2.f32 (pack) + and", out, cyc);
  run<3>("cvt.f32.f16 (unpack)", out, cyc);
  run<4>("fma.f32", out, cyc);
  run<7>("add.f32", out, cyc);
  run<10>("max.f32", out, cyc);
  run<5>("lop3", out, cyc);
  run<6>("prmt", out, cyc);
  run<11>("add.s32", out, cyc);
  run<12>("shf", out, cyc);
  run<13>("fma.f16x2", out, cyc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
