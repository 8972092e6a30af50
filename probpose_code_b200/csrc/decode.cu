// Fused ProbMap decode for sm_100a: one CTA per (person, keypoint) heatmap.
//
//   logits (pass, flipped pass) --/T--> sparsemax --*normalize, clamp[0,1]--> P, Pf
//   merged = 0.5 * (P + mirror(Pf[flip_idx[k]]))                (flip-TTA)
//   C = merged (*) OKS-Gaussian_k, separable, reflect border     (only over the support's
//                                                                 dilated bounding box)
//   (y*, x*) = first arg max C ; one quadratic sub-pixel step on C ; conf = merged[y*, x*]
//   record = [x, y, conf, prob, vis, oks, err / diag]
//
// Reference semantics: probmap_head.py:641-645,757-798, tta.py:35-39,
// post_processing.py:13-39,308-430 (see include/probpose_b200.h: pp_decode).
//
// The kernel is HBM-bound by design: each map (12 KB, 24 KB with TTA) is read exactly
// once with 128-bit streaming loads into registers; everything else lives in shared
// memory; the only global write is the 28-byte record.
#include "common.cuh"

#include <math.h>

namespace pp {

constexpr int kMaxRadius = 9;  // ceil(3 * 3.0): the variance is clipped to <= 3.0
constexpr int kTaps = 2 * kMaxRadius + 1;
constexpr int kDecodeThreads = 256;

struct DecodeParams {
  const float* maps;
  const float* maps_flip;
  const float* scal;
  const float* scal_flip;
  float* records;
  float* merged_out;
  int num_kpts;
  int is_logits;
  float temperature, normalize, err_div;
  int flip_idx[PP_MAX_KEYPOINTS];
  int radius[PP_MAX_KEYPOINTS];
  float taps[PP_MAX_KEYPOINTS][kTaps + 1];  // 1-D factor g of the OKS kernel, sum 1
};

// Block-wide all-reduce of N values through double-buffered scratch: one barrier per call
// (a thread can be at most one reduction ahead of the slowest).
template <int N, typename T, typename Op>
__device__ __forceinline__ void block_allreduce(T (&v)[N], T (*scratch)[kDecodeThreads / 32][4], int& parity, Op op) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] = op(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) scratch[parity][warp][i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T r = scratch[parity][0][i];
#pragma unroll
    for (int w = 1; w < kDecodeThreads / 32; ++w) r = op(r, scratch[parity][w][i]);
    v[i] = r;
  }
  parity ^= 1;
}

__device__ __forceinline__ int reflect(int i, int n) {  // scipy 'reflect': d c b a | a b c d | d c b a
  return i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i);
}

template <int H, int W>
__global__ void __launch_bounds__(kDecodeThreads, 4) decode_kernel(const DecodeParams p) {
  constexpr int NPX = H * W;
  constexpr int NV4 = NPX / 4;
  constexpr int T = kDecodeThreads;
  constexpr int V = NV4 / T;
  static_assert(NV4 % T == 0 && W % 4 == 0, "map must split into whole float4 per thread");

  __shared__ __align__(16) float sP[NPX];  // merged, normalised heatmap
  __shared__ __align__(16) float sH[NPX];  // after the horizontal pass
  __shared__ __align__(16) float sC[NPX];  // convolved map
  __shared__ float red_f[2][T / 32][4];
  __shared__ int red_i[2][T / 32][4];
  int par_f = 0, par_i = 0;

  const int tid = threadIdx.x;
  const int K = p.num_kpts;
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  const bool tta = p.maps_flip != nullptr;
  const int kf = tta ? p.flip_idx[k] : k;

  // ---- 1. stream the map(s) into registers ------------------------------------------------
  float z1[V][4], z2[V][4];
  {
    const float4* s1 = reinterpret_cast<const float4*>(p.maps + (size_t)(b * K + k) * NPX);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 t = ld_stream_f4(s1 + tid + j * T);
      z1[j][0] = t.x; z1[j][1] = t.y; z1[j][2] = t.z; z1[j][3] = t.w;
    }
    if (tta) {
      const float4* s2 = reinterpret_cast<const float4*>(p.maps_flip + (size_t)(b * K + kf) * NPX);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float4 t = ld_stream_f4(s2 + tid + j * T);
        z2[j][0] = t.x; z2[j][1] = t.y; z2[j][2] = t.z; z2[j][3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) z2[j][0] = z2[j][1] = z2[j][2] = z2[j][3] = 0.f;
    }
  }

  // ---- 2. sparsemax(z / T) * normalize, clamp to [0, 1] (both passes together) -----------
  if (p.is_logits) {
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        z1[j][c] = z1[j][c] / p.temperature;
        z2[j][c] = z2[j][c] / p.temperature;
        mx[0] = fmaxf(mx[0], z1[j][c]);
        mx[1] = fmaxf(mx[1], z2[j][c]);
      }
    block_allreduce<2>(mx, red_f, par_f, [](float a, float b2) { return fmaxf(a, b2); });
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) { z1[j][c] -= mx[0]; z2[j][c] -= mx[1]; }

    // Michelot's fixed point on the candidate set {z > -1} (a superset of the support):
    // tau <- (sum_{z > tau} z - 1) / #{z > tau} until the set stops shrinking.
    float thr[2] = {-1.f, -1.f}, tau[2] = {-1.f, -1.f};
    float prev_n[2] = {-1.f, -1.f};
    bool done[2] = {false, !tta};
    for (int it = 0; it < NPX + 2; ++it) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};  // S1, n1, S2, n2 (counts are exact in fp32: <= 3072)
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (z1[j][c] > thr[0]) { acc[0] += z1[j][c]; acc[1] += 1.f; }
          if (z2[j][c] > thr[1]) { acc[2] += z2[j][c]; acc[3] += 1.f; }
        }
      block_allreduce<4>(acc, red_f, par_f, [](float a, float b2) { return a + b2; });
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (done[s]) continue;
        const float n = acc[2 * s + 1];
        tau[s] = (acc[2 * s] - 1.f) / n;  // n >= 1: the maximum (z == 0) is always a candidate
        if (n == prev_n[s]) done[s] = true;
        prev_n[s] = n;
        thr[s] = fmaxf(thr[s], tau[s]);
      }
      if (done[0] && done[1]) break;  // block-uniform: every thread sees the same sums
    }
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        z1[j][c] = fminf(fmaxf(fmaxf(z1[j][c] - tau[0], 0.f) * p.normalize, 0.f), 1.f);
        z2[j][c] = fminf(fmaxf(fmaxf(z2[j][c] - tau[1], 0.f) * p.normalize, 0.f), 1.f);
      }
  }

  // ---- 3. flip-TTA merge into shared memory ----------------------------------------------
#pragma unroll
  for (int j = 0; j < V; ++j)
    reinterpret_cast<float4*>(sP)[tid + j * T] = make_float4(z1[j][0], z1[j][1], z1[j][2], z1[j][3]);
  if (tta) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int v4 = tid + j * T;          // float4 index in the flipped map
      const int y = v4 / (W / 4), xq = v4 % (W / 4);
      const int dst = y * (W / 4) + (W / 4 - 1 - xq);  // mirrored float4 slot, lanes reversed
      float4 a = reinterpret_cast<float4*>(sP)[dst];
      a.x = (a.x + z2[j][3]) * 0.5f;
      a.y = (a.y + z2[j][2]) * 0.5f;
      a.z = (a.z + z2[j][1]) * 0.5f;
      a.w = (a.w + z2[j][0]) * 0.5f;
      reinterpret_cast<float4*>(sP)[dst] = a;
    }
  }
  __syncthreads();

  // ---- 4. support bounding box (any negative value -> treat the map as dense) -------------
  int box[4] = {H, -1, W, -1};  // ymin, ymax, xmin, xmax ; reduce (min, max, min, max) as max of negated
  int neg = 0;
  float4* gout = p.merged_out ? reinterpret_cast<float4*>(p.merged_out + (size_t)(b * K + k) * NPX) : nullptr;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int v4 = tid + j * T;
    const float4 a = reinterpret_cast<const float4*>(sP)[v4];
    if (gout) gout[v4] = a;
    const int y = v4 / (W / 4), x = (v4 % (W / 4)) * 4;
    const float e[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (e[c] != 0.f) {
        box[0] = min(box[0], y); box[1] = max(box[1], y);
        box[2] = min(box[2], x + c); box[3] = max(box[3], x + c);
      }
      if (e[c] < 0.f) neg = 1;
    }
  }
  {
    int r[4] = {-box[0], box[1], -box[2], box[3]};
    block_allreduce<4>(r, red_i, par_i, [](int a, int b2) { return max(a, b2); });
    int ng[1] = {neg};
    block_allreduce<1>(ng, red_i, par_i, [](int a, int b2) { return max(a, b2); });
    box[0] = -r[0]; box[1] = r[1]; box[2] = -r[2]; box[3] = r[3];
    if (ng[0]) { box[0] = 0; box[1] = H - 1; box[2] = 0; box[3] = W - 1; }
  }

  // ---- 5. separable OKS convolution over the dilated box, argmax --------------------------
  const int rad = p.radius[k];
  const float* g = p.taps[k];
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  int X0 = 0, X1 = -1, Y0 = 0, Y1 = -1;
  const bool nonempty = box[1] >= box[0];
  if (nonempty) {
    X0 = max(0, box[2] - rad); X1 = min(W - 1, box[3] + rad);
    Y0 = max(0, box[0] - rad); Y1 = min(H - 1, box[1] + rad);
    const int wx = X1 - X0 + 1;
    const int nh = (box[1] - box[0] + 1) * wx;
    for (int i = tid; i < nh; i += T) {
      const int y = box[0] + i / wx, x = X0 + i % wx;
      const float* row = sP + y * W;
      float acc = 0.f;
      for (int d = -rad; d <= rad; ++d) acc = fmaf(g[d + rad], row[reflect(x + d, W)], acc);
      sH[y * W + x] = acc;
    }
    __syncthreads();
    const int nv = (Y1 - Y0 + 1) * wx;
    for (int i = tid; i < nv; i += T) {
      const int y = Y0 + i / wx, x = X0 + i % wx;
      float acc = 0.f;
      for (int d = -rad; d <= rad; ++d) {
        const int yy = reflect(y + d, H);
        if (yy >= box[0] && yy <= box[1]) acc = fmaf(g[d + rad], sH[yy * W + x], acc);
      }
      sC[y * W + x] = acc;
      if (acc > best) { best = acc; best_i = y * W + x; }  // i ascending => first max per thread
    }
  }
  // block arg-max, smallest flat index wins ties (np.argmax semantics)
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red_f[par_f][warp][0] = best; red_i[par_i][warp][0] = best_i; }
    __syncthreads();  // also orders the sC writes before thread 0 reads them
  }

  // ---- 6. sub-pixel step + record ----------------------------------------------------------
  if (tid == 0) {
    for (int w = 0; w < T / 32; ++w) {
      const float ov = red_f[par_f][w][0];
      const int oi = red_i[par_i][w][0];
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    // Outside the dilated box C is exactly 0.  In sparse mode all values are >= 0, so a
    // non-positive maximum means C == 0 everywhere and flat index 0 is the first maximum.
    const bool dense = (box[0] == 0 && box[1] == H - 1 && box[2] == 0 && box[3] == W - 1);
    if (!nonempty || (!dense && !(best > 0.f))) best_i = 0;
    const int ys = best_i / W, xs = best_i % W;
    auto cval = [&](int y, int x) -> float {
      return (nonempty && y >= Y0 && y <= Y1 && x >= X0 && x <= X1) ? sC[y * W + x] : 0.f;
    };
    float lx = (float)xs, ly = (float)ys;
    if (xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1) {
      const float c = cval(ys, xs), r = cval(ys, xs + 1), l = cval(ys, xs - 1);
      const float dn = cval(ys + 1, xs), up = cval(ys - 1, xs);
      const float dx = __fmul_rn(__fsub_rn(r, l), 0.5f), dy = __fmul_rn(__fsub_rn(dn, up), 0.5f);
      float dxx = __fsub_rn(__fadd_rn(r, l), __fmul_rn(2.f, c));
      float dyy = __fsub_rn(__fadd_rn(dn, up), __fmul_rn(2.f, c));
      if (dxx == 0.f) dxx = 1e-6f;
      if (dyy == 0.f) dyy = 1e-6f;
      lx = __fadd_rn(lx, __fdiv_rn(-dx, dxx));
      ly = __fadd_rn(ly, __fdiv_rn(-dy, dyy));
    }
    float* rec = p.records + (size_t)(b * K + k) * PP_RECORD_FLOATS;
    rec[0] = lx;
    rec[1] = ly;
    rec[2] = sP[best_i];
  }
  if (tid >= 32 && tid < 36) {  // the four scalar branches: prob, vis, oks, err
    const int j = tid - 32;
    float s = 0.f;
    if (p.scal) {
      s = p.scal[(size_t)(b * 4 + j) * K + k];
      if (p.scal_flip) s = (s + p.scal_flip[(size_t)(b * 4 + j) * K + kf]) * 0.5f;
      if (j == 3) s = s / p.err_div;
    }
    p.records[(size_t)(b * K + k) * PP_RECORD_FLOATS + 3 + j] = s;
  }
}

// 1-D factor of the reference's OKS kernel (post_processing.py:13-39), computed in double.
static const double kCocoSigmas[PP_MAX_KEYPOINTS] = {0.026, 0.025, 0.025, 0.035, 0.035, 0.079, 0.079, 0.072, 0.072,
                                                     0.062, 0.062, 0.107, 0.107, 0.087, 0.087, 0.089, 0.089};

static void fill_oks_taps(DecodeParams& p, int K, int H, int W) {
  const double area = sqrt((double)H / 1.25 * (double)W / 1.25);
  for (int k = 0; k < K; ++k) {
    double s = (kCocoSigmas[k] * 2) * (kCocoSigmas[k] * 2) * area * 2;
    s = s < 0.55 ? 0.55 : (s > 3.0 ? 3.0 : s);
    const int r = (int)ceil(s * 3);
    p.radius[k] = r;
    double g[kTaps], sum = 0;
    for (int i = 0; i <= 2 * r; ++i) {
      const double d = i - r;
      g[i] = exp(-(d * d) / (2 * s));
      sum += g[i];
    }
    for (int i = 0; i < kTaps + 1; ++i) p.taps[k][i] = i <= 2 * r ? (float)(g[i] / sum) : 0.f;
  }
}

}  // namespace pp

extern "C" int pp_decode(const pp_decode_cfg* cfg, const float* maps, const float* maps_flip,
                         const int32_t* flip_indices, const float* scalars, const float* scalars_flip,
                         int32_t batch, float* records, float* merged_out, void* stream) {
  using namespace pp;
  PP_REQUIRE(cfg != nullptr, PP_ERR_INVALID, "pp_decode: cfg must be non-NULL");
  PP_REQUIRE(batch >= 0, PP_ERR_INVALID, "pp_decode: negative batch %d", batch);
  PP_REQUIRE(batch == 0 || (maps && records), PP_ERR_INVALID, "pp_decode: maps and records must be non-NULL");
  PP_REQUIRE(cfg->num_keypoints >= 1 && cfg->num_keypoints <= PP_MAX_KEYPOINTS, PP_ERR_INVALID,
             "pp_decode: num_keypoints %d outside [1, %d] (OKS sigma table, post_processing.py:16)",
             cfg->num_keypoints, PP_MAX_KEYPOINTS);
  PP_REQUIRE(cfg->height == 64 && cfg->width == 48, PP_ERR_UNSUPPORTED,
             "pp_decode: heatmap %dx%d not built (only 64x48)", cfg->height, cfg->width);
  PP_REQUIRE(!maps_flip || flip_indices, PP_ERR_INVALID, "pp_decode: maps_flip given without flip_indices");
  PP_REQUIRE(!cfg->input_is_logits || cfg->temperature > 0.f, PP_ERR_INVALID, "pp_decode: temperature must be > 0");
  PP_REQUIRE(!scalars_flip || scalars, PP_ERR_INVALID, "pp_decode: scalars_flip given without scalars");
  if (batch == 0) return PP_OK;

  DecodeParams p;
  p.maps = maps; p.maps_flip = maps_flip; p.scal = scalars; p.scal_flip = scalars_flip;
  p.records = records; p.merged_out = merged_out;
  p.num_kpts = cfg->num_keypoints;
  p.is_logits = cfg->input_is_logits;
  p.temperature = cfg->temperature; p.normalize = cfg->normalize;
  p.err_div = cfg->error_divisor > 0.f
                  ? cfg->error_divisor
                  : sqrtf((float)(cfg->height * cfg->height + cfg->width * cfg->width));
  for (int k = 0; k < PP_MAX_KEYPOINTS; ++k) {
    int f = k;
    if (maps_flip && k < cfg->num_keypoints) {
      f = flip_indices[k];
      PP_REQUIRE(f >= 0 && f < cfg->num_keypoints, PP_ERR_INVALID, "pp_decode: flip_indices[%d]=%d out of range", k, f);
    }
    p.flip_idx[k] = f;
    p.radius[k] = 0;
  }
  fill_oks_taps(p, cfg->num_keypoints, cfg->height, cfg->width);

  const int64_t grid = (int64_t)batch * cfg->num_keypoints;
  PP_REQUIRE(grid < (1ll << 31), PP_ERR_INVALID, "pp_decode: batch too large");
  decode_kernel<64, 48><<<(unsigned)grid, kDecodeThreads, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}
