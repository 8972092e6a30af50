// Fused UDPHeatmap (DARK-UDP) decode for sm_100a - the decode of the ViTPose td-hm configs.
//
//   merged = 0.5 * (H + mirror(Hf[flip_idx[k]]))                       (flip-TTA, heatmap_head.py:245-256, tta.py:35-39)
//   (x, y) = first arg max of merged, score = max ; (-1, -1) if max <= 0  (post_processing.py:178-217)
//   B = zero-padded separable Gaussian blur, bit for bit cv2.GaussianBlur's float arithmetic, rescaled to the old
//       maximum (post_processing.py:220-249)
//   L = log(clip(B, 1e-3, 50)), edge-padded ; gradient / Hessian of L at the peak from 7 samples ;
//   (x, y) -= pinv(Hessian + eps I) grad                                 (refinement.py:102-160)
//   record = [x, y, score]   (heatmap pixels; the caller applies udp_heatmap.py:194-195 in double)
//
// One CTA (8 warps) per (person, keypoint) map: the map is read once from HBM with coalesced 128-bit loads into a
// shared-memory plane (the flipped pass merged in place), block-wide first-max reduction, row pass and column pass of
// the blur through two more planes, block max of the blurred map; one thread then does the 7-sample refinement (the
// Hessian solve in double, exactly where numpy promotes to float64).
#include "common.cuh"

#include <math.h>

namespace pp {

namespace {

constexpr int kUdpThreads = 256;
constexpr int kUdpWarps = kUdpThreads / 32;
constexpr int kUdpMaxTaps = 31;

struct UdpParams {
  const float* maps;
  const float* maps_flip;
  float* records;
  float* merged_out;
  int num_kpts;
  int ntaps;
  int flip_idx[PP_MAX_KEYPOINTS];
  float taps[kUdpMaxTaps];
};

template <int H, int W>
__global__ void __launch_bounds__(kUdpThreads, 3) udp_decode_kernel(const __grid_constant__ UdpParams p) {
  constexpr int NPX = H * W, NV4 = NPX / 4, V = NV4 / kUdpThreads;
  static_assert(NV4 % kUdpThreads == 0 && W % 4 == 0, "map must split into whole float4 per thread");
  __shared__ __align__(16) float sP[NPX];  // merged heatmap
  __shared__ __align__(16) float sR[NPX];  // row pass
  __shared__ __align__(16) float sC[NPX];  // blurred map
  __shared__ float red_v[kUdpWarps];
  __shared__ int red_i[kUdpWarps];
  __shared__ float red_b[kUdpWarps];
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int item = blockIdx.x, K = p.num_kpts;
  const int b = item / K, k = item % K;
  const bool tta = p.maps_flip != nullptr;
  pdl_wait();

  // One map through merge -> first maximum -> blur (sP merged map, sC blurred map); every thread returns the map's
  // maximum, its first flat index and the maximum of the blurred map.
  auto process = [&](int it, bool write_merged, float& best_out, int& best_i_out, float& bmax_out) {
    const int bb = it / K, kk = it % K;
    {
      const float4* s1 = reinterpret_cast<const float4*>(p.maps + (size_t)it * NPX);
#pragma unroll
      for (int j = 0; j < V; ++j) reinterpret_cast<float4*>(sP)[tid + j * kUdpThreads] = ld_stream_f4(s1 + tid + j * kUdpThreads);
      if (tta) {
        const float4* s2 = reinterpret_cast<const float4*>(p.maps_flip + (size_t)(bb * K + p.flip_idx[kk]) * NPX);
        float4 z[V];
#pragma unroll
        for (int j = 0; j < V; ++j) z[j] = ld_stream_f4(s2 + tid + j * kUdpThreads);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int v4 = tid + j * kUdpThreads, y = v4 / (W / 4), xq = v4 % (W / 4);
          const int dst = y * (W / 4) + (W / 4 - 1 - xq);  // mirrored float4 slot, components reversed
          float4 a = reinterpret_cast<float4*>(sP)[dst];
          a.x = (a.x + z[j].w) * 0.5f; a.y = (a.y + z[j].z) * 0.5f; a.z = (a.z + z[j].y) * 0.5f; a.w = (a.w + z[j].x) * 0.5f;
          reinterpret_cast<float4*>(sP)[dst] = a;
        }
      }
    }
    __syncthreads();

    // ---- first maximum of the merged map (smallest flat index among equal maxima) ----
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    float4* gout = (write_merged && p.merged_out) ? reinterpret_cast<float4*>(p.merged_out + (size_t)it * NPX) : nullptr;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int v4 = tid + j * kUdpThreads;
      const float4 a = reinterpret_cast<const float4*>(sP)[v4];
      if (gout) gout[v4] = a;
      const float e[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (e[c] > best) { best = e[c]; best_i = 4 * v4 + c; }  // ascending index per thread
    }
    // NaN-free inputs assumed (np.argmax would return the first NaN)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red_v[warp] = best; red_i[warp] = best_i; }

    // ---- the blur, bit for bit as cv2.GaussianBlur computes it for CV_32F (its AVX2 / FMA3 filter engine; verified
    // against cv2 on the host, see oracle/udp_oracle.py:gaussian_blur_exact): the ROW filter is the general form -
    // s = x[0] k[0], then s = fma(x[j], k[j], s) for j ascending - and the COLUMN filter the symmetric form -
    // s = r[c] k[c], then s = fma(r[c + d] + r[c - d], k[c + d], s) for d = 1 .. radius.  Zero outside the map (the
    // reference pads the map with a radius-wide zero frame before it calls OpenCV, so the border mode never matters).
    // Register-blocked: a thread owns RB consecutive outputs of one row and slides over RB + ntaps - 1 inputs held in
    // registers - one shared-memory read per input, not per tap ----
    const int rad = p.ntaps >> 1;
    constexpr int RB = 12;                       // W / RB segments per row
    static_assert(W % RB == 0 && (H * (W / RB)) % kUdpThreads == 0, "row segments must tile the block");
    constexpr int kWin = RB + kUdpMaxTaps - 1;
    for (int sgm = tid; sgm < H * (W / RB); sgm += kUdpThreads) {
      const int y = sgm / (W / RB), x0 = (sgm % (W / RB)) * RB;
      const float* row = sP + y * W;
      float win[kWin];
#pragma unroll
      for (int i = 0; i < kWin; ++i) {
        const int xx = x0 - rad + i;
        win[i] = (i < RB + 2 * rad && xx >= 0 && xx < W) ? row[xx] : 0.f;
      }
      float acc[RB];
#pragma unroll
      for (int o = 0; o < RB; ++o) acc[o] = 0.f;
#pragma unroll
      for (int j = 0; j < kUdpMaxTaps; ++j) {
        if (j < p.ntaps) {  // block-uniform
          const float t = p.taps[j];
#pragma unroll
          for (int o = 0; o < RB; ++o) acc[o] = fmaf(win[o + j], t, acc[o]);  // j = 0: fma(x, k, 0) = the rounded product
        }
      }
#pragma unroll
      for (int o = 0; o < RB; ++o) sR[y * W + x0 + o] = acc[o];
    }
    __syncthreads();
    // ---- column pass, maximum of the blurred map: a thread owns CB consecutive rows of one column ----
    float bmax = -INFINITY;
    constexpr int CB = 16;
    static_assert(H % CB == 0, "column segments");
    constexpr int kWinC = CB + kUdpMaxTaps - 1;
    constexpr int kMaxRad = kUdpMaxTaps / 2;
    for (int sgm = tid; sgm < W * (H / CB); sgm += kUdpThreads) {
      const int x = sgm % W, y0 = (sgm / W) * CB;  // consecutive threads -> consecutive columns: conflict-free
      // the window is centred at a compile-time position (tap radius kMaxRad) so that win[] stays in registers
      float win[kWinC];
#pragma unroll
      for (int i = 0; i < kWinC; ++i) {
        const int yy = y0 - kMaxRad + i;
        win[i] = (i >= kMaxRad - rad && i < CB + kMaxRad + rad && yy >= 0 && yy < H) ? sR[yy * W + x] : 0.f;
      }
      float acc[CB];
      {
        const float tc = p.taps[rad];
#pragma unroll
        for (int o = 0; o < CB; ++o) acc[o] = __fmul_rn(win[o + kMaxRad], tc);
      }
#pragma unroll
      for (int d = 1; d <= kMaxRad; ++d) {
        if (d <= rad) {  // block-uniform
          const float t = p.taps[rad + d];
#pragma unroll
          for (int o = 0; o < CB; ++o) acc[o] = fmaf(__fadd_rn(win[o + kMaxRad + d], win[o + kMaxRad - d]), t, acc[o]);
        }
      }
#pragma unroll
      for (int o = 0; o < CB; ++o) {
        sC[(y0 + o) * W + x] = acc[o];
        bmax = fmaxf(bmax, acc[o]);
      }
    }
    bmax = warp_max(bmax);
    if (lane == 0) red_b[warp] = bmax;
    __syncthreads();
    for (int w = 0; w < kUdpWarps; ++w) {  // every thread: the same final values
      if (red_v[w] > best || (red_v[w] == best && red_i[w] < best_i)) { best = red_v[w]; best_i = red_i[w]; }
      bmax = fmaxf(bmax, red_b[w]);
    }
    best_out = best; best_i_out = best_i; bmax_out = bmax;
  };
  // log of the clipped, rescaled blurred map in sC, edge-padded (refinement.py:123-127).  The rescale
  // heatmaps[k] *= origin_max / (np.max(heatmaps[k]) + 1e-12) (post_processing.py:247) is float32 arithmetic end to end
  // under NumPy 2 promotion rules (the python float 1e-12 is "weak": float32 + 1e-12 stays float32; this image ships
  // NumPy 2.3, and the oracle is pinned to the reference run under it)
  auto L_at = [&](int y, int x, float scale) -> float {
    y = min(max(y, 0), H - 1);
    x = min(max(x, 0), W - 1);
    const float v = __fmul_rn(sC[y * W + x], scale);
    // the correctly rounded float logarithm (double log, rounded once): numpy's float32 log is correctly rounded for
    // all but a few inputs, CUDA's logf (1 ulp) for fewer - and a flat map's Hessian amplifies every ulp
    return (float)log((double)fminf(fmaxf(v, 1e-3f), 50.0f));
  };

  float best, bmax;
  int best_i;
  process(item, true, best, best_i, bmax);
  const float scale = __fdiv_rn(best, __fadd_rn(bmax, 1e-12f));
  float i_, ix1, iy1, ix1y1, ix1_y1_, ix1_, iy1_;
  float px_f, py_f;
  if (best > 0.f) {
    const int py = best_i / W, px = best_i % W;
    px_f = (float)px; py_f = (float)py;
    i_ = L_at(py, px, scale); ix1 = L_at(py, px + 1, scale); iy1 = L_at(py + 1, px, scale); ix1y1 = L_at(py + 1, px + 1, scale);
    ix1_y1_ = L_at(py - 1, px - 1, scale); ix1_ = L_at(py, px - 1, scale); iy1_ = L_at(py - 1, px, scale);
  } else {
    // No response (maximum <= 0): get_heatmap_maximum marks the keypoint (-1, -1) (post_processing.py:213-215), and the
    // refinement then indexes the FLATTENED edge-padded stack of the K maps at (-1 + 1) + (-1 + 1) (W + 2) = the first
    // element of map k's padded plane (refinement.py:130-138): the four "forward" samples are this map's corner
    // L(0, 0), the three "backward" samples fall off the front of the plane into the END of the previous keypoint's
    // padded plane (keypoint K - 1 for k = 0: negative indices wrap in numpy) - its bottom-right corner, bottom-left
    // corner and bottom-right corner again.  Reproduced as it is: the previous map goes through the same pipeline.
    px_f = -1.f; py_f = -1.f;
    i_ = ix1 = iy1 = ix1y1 = L_at(0, 0, scale);
    __syncthreads();  // every thread is done with this map's planes
    float nbest, nbmax;
    int nbest_i;
    process(b * K + (k + K - 1) % K, false, nbest, nbest_i, nbmax);
    const float nscale = __fdiv_rn(nbest, __fadd_rn(nbmax, 1e-12f));
    ix1_ = L_at(H - 1, W - 1, nscale);      // flat index - 1:       padded (H + 1, W + 1) of the previous map
    iy1_ = L_at(H - 1, 0, nscale);          // flat index - (W + 2): padded (H + 1, 0)
    ix1_y1_ = L_at(H - 1, W - 1, nscale);   // flat index - (W + 3): padded (H, W + 1)
  }
  if (tid == 0) {
    float* rec = p.records + (size_t)item * 3;
    rec[2] = best;
    {
      const float dx = __fmul_rn(0.5f, __fsub_rn(ix1, ix1_)), dy = __fmul_rn(0.5f, __fsub_rn(iy1, iy1_));
      const float dxx = __fadd_rn(__fsub_rn(ix1, __fmul_rn(2.f, i_)), ix1_);
      const float dyy = __fadd_rn(__fsub_rn(iy1, __fmul_rn(2.f, i_)), iy1_);
      float t = __fsub_rn(ix1y1, ix1);
      t = __fsub_rn(t, iy1); t = __fadd_rn(t, i_); t = __fadd_rn(t, i_); t = __fsub_rn(t, ix1_); t = __fsub_rn(t, iy1_);
      t = __fadd_rn(t, ix1_y1_);
      const float dxy = __fmul_rn(0.5f, t);
      // float64 from here on, like numpy (the eps * eye(2) term is float64): pinv of a symmetric 2 x 2
      const double eps = 1.1920928955078125e-07;
      const double a = (double)dxx + eps, c = (double)dyy + eps, bq = (double)dxy;
      const double half_tr = 0.5 * (a + c), rt = sqrt(0.25 * (a - c) * (a - c) + bq * bq);
      const double l1 = half_tr + rt, l2 = half_tr - rt;
      const double smax = fmax(fabs(l1), fabs(l2)), smin = fmin(fabs(l1), fabs(l2));
      double ox, oy;
      if (smax == 0.0) {
        ox = oy = 0.0;  // pinv(0) = 0
      } else if (smin > 1e-15 * smax) {
        const double det = a * c - bq * bq;
        ox = (c * (double)dx - bq * (double)dy) / det;
        oy = (a * (double)dy - bq * (double)dx) / det;
      } else {
        // rank 1: only the dominant eigen-pair survives numpy's rcond cut-off
        const double l = fabs(l1) >= fabs(l2) ? l1 : l2;
        double vx = bq, vy = l - a;
        if (fabs(vx) + fabs(vy) == 0.0) { vx = l - c; vy = bq; }
        if (fabs(vx) + fabs(vy) == 0.0) { vx = fabs(a) >= fabs(c) ? 1.0 : 0.0; vy = 1.0 - vx; }
        const double n2 = vx * vx + vy * vy, proj = (vx * (double)dx + vy * (double)dy) / (n2 * l);
        ox = vx * proj;
        oy = vy * proj;
      }
      rec[0] = (float)((double)px_f - ox);
      rec[1] = (float)((double)py_f - oy);
    }
  }
}

void fill_gaussian_taps(UdpParams& p, int ksize) {  // OpenCV getGaussianKernel(ksize, sigma <= 0) for CV_32F
  const double sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8;
  double g[kUdpMaxTaps], sum = 0;
  for (int i = 0; i < ksize; ++i) {
    const double x = i - (ksize - 1) * 0.5;
    g[i] = exp(-(x * x) / (2 * sigma * sigma));
    sum += g[i];
  }
  for (int i = 0; i < kUdpMaxTaps; ++i) p.taps[i] = i < ksize ? (float)(g[i] / sum) : 0.f;
  p.ntaps = ksize;
}

}  // namespace

}  // namespace pp

extern "C" int pp_decode_udp(const pp_udp_cfg* cfg, const float* maps, const float* maps_flip, const int32_t* flip_indices,
                             int32_t batch, float* records, float* merged_out, void* stream) {
  using namespace pp;
  PP_REQUIRE(cfg != nullptr, PP_ERR_INVALID, "pp_decode_udp: cfg must be non-NULL");
  PP_REQUIRE(batch >= 0, PP_ERR_INVALID, "pp_decode_udp: negative batch %d", batch);
  PP_REQUIRE(batch == 0 || (maps && records), PP_ERR_INVALID, "pp_decode_udp: maps and records must be non-NULL");
  PP_REQUIRE(cfg->num_keypoints >= 1 && cfg->num_keypoints <= PP_MAX_KEYPOINTS, PP_ERR_INVALID,
             "pp_decode_udp: num_keypoints %d outside [1, %d]", cfg->num_keypoints, PP_MAX_KEYPOINTS);
  PP_REQUIRE(cfg->height == 64 && cfg->width == 48, PP_ERR_UNSUPPORTED, "pp_decode_udp: heatmap %dx%d not built (only 64x48)",
             cfg->height, cfg->width);
  PP_REQUIRE(cfg->blur_kernel_size >= 3 && cfg->blur_kernel_size <= kUdpMaxTaps && (cfg->blur_kernel_size & 1) == 1, PP_ERR_INVALID,
             "pp_decode_udp: blur_kernel_size %d must be odd and in [3, %d] (post_processing.py:237)", cfg->blur_kernel_size, kUdpMaxTaps);
  PP_REQUIRE(!maps_flip || flip_indices, PP_ERR_INVALID, "pp_decode_udp: maps_flip given without flip_indices");
  if (batch == 0) return PP_OK;
  UdpParams p;
  p.maps = maps; p.maps_flip = maps_flip; p.records = records; p.merged_out = merged_out;
  p.num_kpts = cfg->num_keypoints;
  for (int k = 0; k < PP_MAX_KEYPOINTS; ++k) {
    int f = k;
    if (maps_flip && k < cfg->num_keypoints) {
      f = flip_indices[k];
      PP_REQUIRE(f >= 0 && f < cfg->num_keypoints, PP_ERR_INVALID, "pp_decode_udp: flip_indices[%d]=%d out of range", k, f);
    }
    p.flip_idx[k] = f;
  }
  fill_gaussian_taps(p, cfg->blur_kernel_size);
  const int64_t count = (int64_t)batch * cfg->num_keypoints;
  PP_REQUIRE(count < (1ll << 31), PP_ERR_INVALID, "pp_decode_udp: batch too large");
  PP_CHECK_CUDA(launch_pdl(udp_decode_kernel<64, 48>, dim3((unsigned)count), dim3(kUdpThreads), 0, (cudaStream_t)stream, p));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}
