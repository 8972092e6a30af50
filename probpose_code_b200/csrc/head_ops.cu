// Non-GEMM kernels of ProbMapHead: tap gathers for the conv / deconv GEMMs, MaxPool + ReLU,
// the tail of the four scalar branches, and weight packing (deconv phases, 3x3 taps, BN fold).
//
// Reference semantics: mmpose/models/heads/hybrid_heads/probmap_head.py:261-410 (branches:
// Conv3x3 -> BN -> MaxPool -> ReLU, x3, then Conv1x1 -> Sigmoid / ReLU), :435-472 (deconv
// k4 s2 p1 + BN + ReLU), :244-247 (final 1x1 conv).
#include "engine_ops.cuh"

namespace pp {

// ---- tap gather ---------------------------------------------------------------------------
// Works on 16-byte chunks of the operand planes (FP16X3 rows are [hi plane | lo plane]).
__global__ void __launch_bounds__(256) gather_taps_kernel(const GatherParams p, int planes, int eb, const uint8_t* src,
                                                          uint8_t* dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int cpc = p.c * eb / 16;  // chunks per (tap, plane)
  const size_t src_row_bytes = (size_t)planes * p.src_c * eb;
  const size_t dst_plane_bytes = (size_t)p.ntaps * p.c * eb;
  // one warp per (output row, plane, tap): the index arithmetic once per warp, the lanes copy the tap's contiguous
  // c * eb bytes in 16-byte pieces (a thread per piece spent ~150 instructions of divisions on every 16 bytes)
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t items = (int64_t)p.batch * p.h * p.w * planes * p.ntaps;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += warps) {
    const int t = (int)(it % p.ntaps);
    int64_t r = it / p.ntaps;
    const int pl = (int)(r % planes); r /= planes;
    const int x = (int)(r % p.w);
    const int y = (int)((r / p.w) % p.h);
    const int b = (int)(r / ((int64_t)p.w * p.h));
    const int sy = y + p.dy[t], sx = x + p.dx[t];
    const bool inside = sy >= 0 && sy < p.h && sx >= 0 && sx < p.w;
    const size_t srow = ((size_t)b * p.h + sy) * p.w + sx;
    const uint8_t* s = src + srow * src_row_bytes + ((size_t)pl * p.src_c + p.c_off) * eb;
    uint8_t* d = dst + (size_t)r * planes * dst_plane_bytes + pl * dst_plane_bytes + ((size_t)t * p.c) * eb;
    for (int ch = lane; ch < cpc; ch += 32) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (inside) v = *reinterpret_cast<const uint4*>(s + (size_t)ch * 16);
      *reinterpret_cast<uint4*>(d + (size_t)ch * 16) = v;
    }
  }
}

int launch_gather_taps(int prec, const GatherParams& p, const void* src_op, void* dst_op, cudaStream_t st) {
  const int eb = operand_elem_bytes(prec), planes = prec == PP_PREC_FP16X3 ? 2 : 1;
  const int epc = 16 / eb;
  PP_REQUIRE(p.c % epc == 0 && p.c_off % epc == 0 && p.src_c % epc == 0, PP_ERR_UNSUPPORTED,
             "gather: channels %d/%d/%d not multiples of %d", p.c, p.c_off, p.src_c, epc);
  PP_REQUIRE(p.ntaps >= 1 && p.ntaps <= 9, PP_ERR_INVALID, "gather: %d taps", p.ntaps);
  const int64_t total = (int64_t)p.batch * p.h * p.w * planes * p.ntaps * 32;  // a warp per (row, plane, tap)
  if (total == 0) return PP_OK;
  const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  PP_CHECK_CUDA(launch_pdl(gather_taps_kernel, dim3(grid), dim3(256), 0, st, p, planes, eb, reinterpret_cast<const uint8_t*>(src_op),
                           reinterpret_cast<uint8_t*>(dst_op)));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

// ---- MaxPool + ReLU -----------------------------------------------------------------------
template <int PREC>
__global__ void __launch_bounds__(256) pool_relu_kernel(const float* __restrict__ x, int batch, int h, int w, int c, int ph,
                                                        int pw, void* out_op) {
  pdl_launch_dependents();
  pdl_wait();
  const int oh = h / ph, ow = w / pw, c4 = c / 4;
  const int64_t total = (int64_t)batch * oh * ow * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c4) * 4;
    const int64_t orow = i / c4;
    const int ox = (int)(orow % ow), oy = (int)((orow / ow) % oh), b = (int)(orow / ((int64_t)ow * oh));
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dy = 0; dy < ph; ++dy)
      for (int dx = 0; dx < pw; ++dx) {
        const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * h + oy * ph + dy) * w + ox * pw + dx) * c + ch);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); m.z = fmaxf(m.z, 0.f); m.w = fmaxf(m.w, 0.f);
    store_operand4<PREC>(out_op, orow, ch, c, m);
  }
}

int launch_pool_relu(int prec, const float* x, int batch, int h, int w, int c, int ph, int pw, void* out_op,
                     cudaStream_t st) {
  PP_REQUIRE(h % ph == 0 && w % pw == 0 && c % 4 == 0, PP_ERR_UNSUPPORTED, "pool: %dx%d map, %dx%d window, %d channels", h,
             w, ph, pw, c);
  const int64_t total = (int64_t)batch * (h / ph) * (w / pw) * (c / 4);
  if (total == 0) return PP_OK;
  const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  { cudaError_t lerr = cudaSuccess; PP_DISPATCH_PREC(prec, (lerr = launch_pdl(pool_relu_kernel<PREC>, dim3(grid), dim3(256), 0, st, x, batch, h, w, c, ph, pw, out_op))); PP_CHECK_CUDA(lerr); }
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

// ---- branch tail --------------------------------------------------------------------------
// One CTA per person: pooled = relu(max over the 2x2 map) for all 4 * c channels in shared
// memory, then one warp per output (branch, keypoint) does the c-long dot product.
__global__ void __launch_bounds__(256) branch_tail_kernel(const float* __restrict__ x, int c, int k,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* scalars) {
  extern __shared__ float pooled[];  // 4 * c
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, C4 = 4 * c;
  const float* xb = x + (size_t)b * 4 * C4;
  for (int i = threadIdx.x; i < C4; i += blockDim.x)
    pooled[i] = fmaxf(fmaxf(fmaxf(xb[i], xb[C4 + i]), fmaxf(xb[2 * C4 + i], xb[3 * C4 + i])), 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = warp; o < 4 * k; o += nw) {
    const int br = o / k;
    const float* wr = w + (size_t)o * c;
    float acc = 0.f;
    for (int i = lane; i < c; i += 32) acc = fmaf(pooled[br * c + i], wr[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float z = acc + bias[o];
      scalars[(size_t)b * 4 * k + o] = br == 3 ? fmaxf(z, 0.f) : 1.0f / (1.0f + expf(-z));
    }
  }
}

int launch_branch_tail(const float* x, int batch, int c, int k, const float* w, const float* bias, float* scalars,
                       cudaStream_t st) {
  if (batch == 0) return PP_OK;
  PP_CHECK_CUDA(launch_pdl(branch_tail_kernel, dim3(batch), dim3(256), (size_t)4 * c * sizeof(float), st, x, c, k, w, bias, scalars));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

// ---- weight packing -----------------------------------------------------------------------
__global__ void pack_deconv_phase_kernel(const float* __restrict__ w, int cin, int cout, int ky0, int ky1, int kx0, int kx1,
                                         float* out) {
  const int64_t total = (int64_t)cout * 4 * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int t = (int)((i / cin) % 4);
    const int co = (int)(i / (4 * (int64_t)cin));
    const int ky = (t >> 1) ? ky1 : ky0, kx = (t & 1) ? kx1 : kx0;
    out[i] = w[(((size_t)ci * cout + co) * 4 + ky) * 4 + kx];
  }
}

int launch_pack_deconv_phase(const float* w, int cin, int cout, int py, int px, float* out, cudaStream_t st) {
  int d, ky0, ky1, kx0, kx1;
  deconv_tap(py, 0, &d, &ky0); deconv_tap(py, 1, &d, &ky1);
  deconv_tap(px, 0, &d, &kx0); deconv_tap(px, 1, &d, &kx1);
  pack_deconv_phase_kernel<<<148 * 4, 256, 0, st>>>(w, cin, cout, ky0, ky1, kx0, kx1, out);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

__global__ void pack_conv3x3_kernel(const float* __restrict__ w, int cout, int cin, float* out) {
  const int64_t total = (int64_t)cout * 9 * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int t = (int)((i / cin) % 9);
    const int co = (int)(i / (9 * (int64_t)cin));
    out[i] = w[((size_t)co * cin + ci) * 9 + t];
  }
}

int launch_pack_conv3x3(const float* w, int cout, int cin, float* out, cudaStream_t st) {
  pack_conv3x3_kernel<<<148 * 4, 256, 0, st>>>(w, cout, cin, out);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

__global__ void fold_bn_kernel(const float* gamma, const float* beta, const float* mean, const float* var,
                               const float* conv_bias, float eps, int n, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // same operation order as ATen's eval-mode batch_norm: (x - mean) * (gamma / sqrt(var + eps)) + beta
  const float s = gamma[i] / sqrtf(var[i] + eps);
  scale[i] = s;
  shift[i] = ((conv_bias ? conv_bias[i] : 0.f) - mean[i]) * s + beta[i];
}

int launch_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, const float* conv_bias,
                   float eps, int n, float* scale, float* shift, cudaStream_t st) {
  fold_bn_kernel<<<(n + 255) / 256, 256, 0, st>>>(gamma, beta, mean, var, conv_bias, eps, n, scale, shift);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}

}  // namespace pp
