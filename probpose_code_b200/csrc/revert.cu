// Heatmap read-back: per-person inverse warp of the (K, H, W) heatmaps into the (padded) image and element-wise max
// over the persons, in ONE pass over the output.
//
// Replaces, per person on the host in the reference (mmpose/structures/utils.py):
//   revert_heatmap (:146-175): cv2.warpAffine(heatmap_HWK, warp_mat, (img_w, img_h), flags=cv2.INTER_LINEAR)
//   merge_data_samples (:117): np.max(padded_heatmaps, axis=0)
// i.e. P full-image float tensors (141 MB each at 1080p, K = 17) written and re-read on the CPU, by one kernel that
// writes the merged (K, img_h, img_w) tensor once; the P x 208 KB heatmaps stay in L2.
// Per pixel it is OpenCV's warpAffine for CV_32F (imgwarp.cpp): the matrix inverted in double, AB_BITS = 10 fixed-point
// source coordinates rounded like cvRound, 1/32-pixel fractions, float bilinear weights (1 - fx)(1 - fy) ... from the
// 32 x 32 table (exact dyadic rationals), v0 w0 + v1 w1 + v2 w2 + v3 w3 in float in that order, BORDER_CONSTANT 0.
#include "common.cuh"

#include <math.h>

namespace pp {

namespace {

constexpr int kRevThreads = 128;
constexpr int kRevMaxK = PP_MAX_KEYPOINTS;

// inverse matrices (dst -> src), computed once per person by a tiny kernel exactly like cv::warpAffine does
// plus, per person, the output rectangle outside which every sample is the border value (the heatmap's corners
// (-1 .. W, -1 .. H) mapped forward, 2 px of margin): pixels outside skip the exact coordinate arithmetic
constexpr int kRevScratch = 10;  // doubles per person: 6 matrix entries + x0, x1, y0, y1
__global__ void revert_invert_kernel(const double* __restrict__ mats, int n, int hh, int hw, double* __restrict__ inv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double m[6];
  for (int i = 0; i < 6; ++i) m[i] = mats[p * 6 + i];
  {
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int c = 0; c < 4; ++c) {
      const double hx = (c & 1) ? (double)hw : -1.0, hy = (c & 2) ? (double)hh : -1.0;
      const double ix = m[0] * hx + m[1] * hy + m[2], iy = m[3] * hx + m[4] * hy + m[5];
      x0 = fmin(x0, ix); x1 = fmax(x1, ix); y0 = fmin(y0, iy); y1 = fmax(y1, iy);
    }
    const bool finite = isfinite(x0) && isfinite(x1) && isfinite(y0) && isfinite(y1) && (m[0] * m[4] - m[1] * m[3]) != 0.0;
    inv[p * kRevScratch + 6] = finite ? floor(x0) - 2.0 : -1e300;  // degenerate matrix: no rejection
    inv[p * kRevScratch + 7] = finite ? ceil(x1) + 2.0 : 1e300;
    inv[p * kRevScratch + 8] = finite ? floor(y0) - 2.0 : -1e300;
    inv[p * kRevScratch + 9] = finite ? ceil(y1) + 2.0 : 1e300;
  }
  double d = __dsub_rn(__dmul_rn(m[0], m[4]), __dmul_rn(m[1], m[3]));
  d = d != 0.0 ? __ddiv_rn(1.0, d) : 0.0;
  const double a11 = __dmul_rn(m[4], d), a22 = __dmul_rn(m[0], d);
  m[0] = a11; m[1] = __dmul_rn(m[1], -d); m[3] = __dmul_rn(m[3], -d); m[4] = a22;
  const double b1 = __dsub_rn(__dmul_rn(-m[0], m[2]), __dmul_rn(m[1], m[5]));
  const double b2 = __dsub_rn(__dmul_rn(-m[3], m[2]), __dmul_rn(m[4], m[5]));
  m[2] = b1; m[5] = b2;
  for (int i = 0; i < 6; ++i) inv[p * kRevScratch + i] = m[i];
}

template <int K>
__global__ void __launch_bounds__(kRevThreads) revert_merge_kernel(const float* __restrict__ hms, const double* __restrict__ inv, int n,
                                                                  int hh, int hw, float* __restrict__ out, int img_h, int img_w) {
  pdl_wait();
  const int y = blockIdx.y, x = blockIdx.x * kRevThreads + threadIdx.x;
  if (x >= img_w) return;
  float best[K];
#pragma unroll
  for (int k = 0; k < K; ++k) best[k] = -INFINITY;
  const size_t plane = (size_t)hh * hw;
  for (int p = 0; p < n; ++p) {
    const double* m = inv + p * kRevScratch;  // uniform loads
    if ((double)x < m[6] || (double)x > m[7] || (double)y < m[8] || (double)y > m[9]) {  // far outside the footprint
#pragma unroll
      for (int k = 0; k < K; ++k) best[k] = fmaxf(best[k], 0.f);
      continue;
    }
    const int x0i = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
    const int y0i = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
    const int xq = (x0i + adelta) >> 5, yq = (y0i + bdelta) >> 5;
    const int sx = min(max(xq >> 5, -32768), 32767), sy = min(max(yq >> 5, -32768), 32767);
    const bool x0ok = sx >= 0 && sx < hw, x1ok = sx + 1 >= 0 && sx + 1 < hw;
    const bool y0ok = sy >= 0 && sy < hh, y1ok = sy + 1 >= 0 && sy + 1 < hh;
    if (!((x0ok || x1ok) && (y0ok || y1ok))) {  // outside this person's footprint: the border value
#pragma unroll
      for (int k = 0; k < K; ++k) best[k] = fmaxf(best[k], 0.f);
      continue;
    }
    const float fx = (float)(xq & 31) * (1.f / 32.f), fy = (float)(yq & 31) * (1.f / 32.f);
    const float w0 = __fmul_rn(1.f - fy, 1.f - fx), w1 = __fmul_rn(1.f - fy, fx), w2 = __fmul_rn(fy, 1.f - fx), w3 = __fmul_rn(fy, fx);
    const float* base = hms + (size_t)p * K * plane + (ptrdiff_t)sy * hw + sx;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float* s = base + (size_t)k * plane;
      const float v0 = (y0ok && x0ok) ? __ldg(s) : 0.f, v1 = (y0ok && x1ok) ? __ldg(s + 1) : 0.f;
      const float v2 = (y1ok && x0ok) ? __ldg(s + hw) : 0.f, v3 = (y1ok && x1ok) ? __ldg(s + hw + 1) : 0.f;
      const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1)), __fmul_rn(v2, w2)), __fmul_rn(v3, w3));
      best[k] = fmaxf(best[k], v);
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) out[((size_t)k * img_h + y) * img_w + x] = best[k];
}

}  // namespace

}  // namespace pp

extern "C" int pp_revert_heatmaps(const float* heatmaps, const double* warp_mats, int32_t persons, int32_t num_keypoints,
                                  int32_t height, int32_t width, float* out, int32_t img_h, int32_t img_w, double* scratch,
                                  void* stream) {
  using namespace pp;
  PP_REQUIRE(persons >= 1, PP_ERR_INVALID, "pp_revert_heatmaps: needs at least one person (np.max of an empty list raises), got %d", persons);
  PP_REQUIRE(heatmaps && warp_mats && out && scratch, PP_ERR_INVALID, "pp_revert_heatmaps: NULL pointer");
  PP_REQUIRE(num_keypoints == kRevMaxK, PP_ERR_UNSUPPORTED, "pp_revert_heatmaps: built for %d keypoints (got %d)", kRevMaxK, num_keypoints);
  PP_REQUIRE(height > 0 && width > 0 && height < 32768 && width < 32768, PP_ERR_INVALID, "pp_revert_heatmaps: bad heatmap %dx%d", height, width);
  PP_REQUIRE(img_h > 0 && img_w > 0 && img_h <= 65535, PP_ERR_INVALID, "pp_revert_heatmaps: bad image %dx%d", img_h, img_w);
  cudaStream_t st = (cudaStream_t)stream;
  revert_invert_kernel<<<(persons + 63) / 64, 64, 0, st>>>(warp_mats, persons, height, width, scratch);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  const dim3 grid((img_w + kRevThreads - 1) / kRevThreads, img_h);
  PP_CHECK_CUDA(launch_pdl(revert_merge_kernel<kRevMaxK>, grid, dim3(kRevThreads), 0, st, heatmaps, (const double*)scratch, (int)persons,
                           (int)height, (int)width, out, (int)img_h, (int)img_w));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}
