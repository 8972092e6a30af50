// GPU crop front-end: frame + per-person affine matrices -> uint8 model inputs.
//
// Replaces, per person and on the host in the reference, TopdownAffine's
//   cv2.warpAffine(img, warp_mat, (w, h), flags=cv2.INTER_LINEAR)
// (mmpose/datasets/transforms/topdown_transforms.py:126) followed by PackPoseInputs' HWC -> CHW
// (datasets/transforms/formatting.py), bit for bit: the inverse matrix in double, AB_BITS = 10
// fixed-point source coordinates rounded like cvRound, 1/32-pixel fractions, 15-bit bilinear
// weights, (sum + 2^14) >> 15, BORDER_CONSTANT 0 - OpenCV's imgwarp.cpp algorithm.
// The matrices (get_udp_warp_matrix, float32) are computed on the host by the plugin layer exactly
// as the reference computes them; everything per pixel happens here.
#include "common.cuh"

namespace pp {

struct CropParams {
  const uint8_t* frame;  // (fh, fw, 3) BGR, row pitch frame_row_bytes
  int fh, fw;
  int64_t frame_row_bytes;
  const float* mats;     // device (n, 2, 3) forward matrices (source -> crop)
  uint8_t* crops;        // device (n, 3, oh, ow)
  int n, oh, ow;
};

__global__ void __launch_bounds__(256) crop_warp_kernel(const CropParams p) {
  __shared__ double inv[6];
  const int y = blockIdx.x, person = blockIdx.y;
  if (threadIdx.x == 0) {
    // cv::invertAffineTransform as warpAffine does it (no FMA contraction: same roundings as the host code)
    double m[6];
    for (int i = 0; i < 6; ++i) m[i] = (double)p.mats[person * 6 + i];
    double d = __dsub_rn(__dmul_rn(m[0], m[4]), __dmul_rn(m[1], m[3]));
    d = d != 0.0 ? __ddiv_rn(1.0, d) : 0.0;
    const double a11 = __dmul_rn(m[4], d), a22 = __dmul_rn(m[0], d);
    m[0] = a11; m[1] = __dmul_rn(m[1], -d); m[3] = __dmul_rn(m[3], -d); m[4] = a22;
    const double b1 = __dsub_rn(__dmul_rn(-m[0], m[2]), __dmul_rn(m[1], m[5]));
    const double b2 = __dsub_rn(__dmul_rn(-m[3], m[2]), __dmul_rn(m[4], m[5]));
    m[2] = b1; m[5] = b2;
    for (int i = 0; i < 6; ++i) inv[i] = m[i];
  }
  __syncthreads();
  const int x0i = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(inv[1], (double)y), inv[2]), 1024.0)) + 16;
  const int y0i = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(inv[4], (double)y), inv[5]), 1024.0)) + 16;
  for (int x = threadIdx.x; x < p.ow; x += blockDim.x) {
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(inv[0], (double)x), 1024.0));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(inv[3], (double)x), 1024.0));
    const int xq = (x0i + adelta) >> 5, yq = (y0i + bdelta) >> 5;
    const int sx = min(max(xq >> 5, -32768), 32767), sy = min(max(yq >> 5, -32768), 32767);
    const int fx = xq & 31, fy = yq & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    const bool x0ok = sx >= 0 && sx < p.fw, x1ok = sx + 1 >= 0 && sx + 1 < p.fw;
    const bool y0ok = sy >= 0 && sy < p.fh, y1ok = sy + 1 >= 0 && sy + 1 < p.fh;
    const uint8_t* r0 = p.frame + (int64_t)sy * p.frame_row_bytes + (int64_t)sx * 3;
    const uint8_t* r1 = r0 + p.frame_row_bytes;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int acc = 1 << 14;
      if (y0ok && x0ok) acc += w00 * r0[c];
      if (y0ok && x1ok) acc += w01 * r0[3 + c];
      if (y1ok && x0ok) acc += w10 * r1[c];
      if (y1ok && x1ok) acc += w11 * r1[3 + c];
      p.crops[(((int64_t)person * 3 + c) * p.oh + y) * p.ow + x] = (uint8_t)(acc >> 15);
    }
  }
}

}  // namespace pp

extern "C" int pp_crop_warp(const uint8_t* frame_hwc_bgr, int32_t frame_h, int32_t frame_w, int64_t frame_row_bytes,
                            const float* warp_mats, int32_t n, uint8_t* crops, int32_t out_h, int32_t out_w, void* stream) {
  using namespace pp;
  PP_REQUIRE(n >= 0, PP_ERR_INVALID, "pp_crop_warp: negative count %d", n);
  if (n == 0) return PP_OK;
  PP_REQUIRE(frame_hwc_bgr && warp_mats && crops, PP_ERR_INVALID, "pp_crop_warp: NULL pointer");
  PP_REQUIRE(frame_h > 0 && frame_w > 0 && frame_h < 32768 && frame_w < 32768 && frame_row_bytes >= (int64_t)frame_w * 3, PP_ERR_INVALID,
             "pp_crop_warp: bad frame %dx%d pitch %lld (cv2.warpAffine addresses sources with 16-bit coordinates)", frame_h,
             frame_w, (long long)frame_row_bytes);
  PP_REQUIRE(out_h > 0 && out_w > 0 && out_h <= 65535 && n <= 65535, PP_ERR_INVALID, "pp_crop_warp: bad output %dx%d x %d", out_h,
             out_w, n);
  CropParams p;
  p.frame = frame_hwc_bgr; p.fh = frame_h; p.fw = frame_w; p.frame_row_bytes = frame_row_bytes;
  p.mats = warp_mats; p.crops = crops; p.n = n; p.oh = out_h; p.ow = out_w;
  const int threads = out_w >= 256 ? 256 : ((out_w + 31) / 32) * 32;
  crop_warp_kernel<<<dim3(out_h, n), threads, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}
