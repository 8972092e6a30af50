// Fused ProbMap decode for sm_100a.
//
//   logits (pass, flipped pass) --/T--> sparsemax --*normalize, clamp[0,1]--> P, Pf
//   merged = 0.5 * (P + mirror(Pf[flip_idx[k]]))                (flip-TTA)
//   C = merged (*) OKS-Gaussian_k, separable, reflect border
//   (y*, x*) = first arg max C ; one quadratic sub-pixel step on C ; conf = merged[y*, x*]
//   record = [x, y, conf, prob, vis, oks, err / diag]
//
// Reference semantics: probmap_head.py:641-645,757-798, tta.py:35-39,
// post_processing.py:13-39,308-430 (see include/probpose_b200.h: pp_decode).
//
// HBM-bound by design: every map (12 KB, 24 KB with TTA) is read exactly once from HBM; the only
// global write is the 28-byte record.  ONE WARP PER MAP, 8 maps per CTA, 2 CTAs (16 warps) per SM; no map
// is staged in shared memory:
//   * the warp streams its map through registers (24 coalesced 128-bit loads per lane, issued in
//     chunks of 6; only the maximum of each float4 is kept),
//   * max (CREDUX) -> candidate set {z > max - T}: the few float4 that hold a candidate are re-read
//     (L1 / L2 hits) and ballot-compacted into a short list (a trained head leaves 2-3 pixels, see
//     SURVEY.md 8c) -> Michelot's fixed point on the list = the exact sort / cumsum threshold of
//     sparsemax,
//   * the convolution is evaluated sparsely, C(q) = sum_s w_s g(dy) g(dx) (+ reflected images), only
//     where the arg max can be - inside the support's bounding box (the OKS kernel is non-negative and
//     decreasing, so the maximum of C cannot lie outside it): directly at the pixels of a compact box,
//     or - scattered supports - by scattering every source's window into the warp's
//     64 x 48 shared-memory tile and scanning the rows of the box,
//   * all reductions are warp shuffles / redux: no block barrier on this path.
// Maps that are not sparse (flat random-init logits, arbitrary / negative heatmaps handed to the
// public codec API) are queued and decoded afterwards by the whole CTA with a dense separable
// convolution (decode_dense) - correct for any input, just not HBM-bound.
#include "common.cuh"

#include <math.h>

namespace pp {

constexpr int kMaxRadius = 9;  // ceil(3 * 3.0): the variance is clipped to <= 3.0
constexpr int kTaps = 2 * kMaxRadius + 1;
constexpr int kDecWarps = 8;    // maps per CTA (one warp each)
constexpr int kDecThreads = 32 * kDecWarps;
constexpr int kDecCtasPerSm = 2;  // 16 maps in flight per SM
constexpr int kDenseWarps = kDecWarps;   // the dense path uses the whole CTA (3 float4 of the map per thread)
constexpr int kDenseThreads = 32 * kDenseWarps;
constexpr int kListCap = 128;   // candidates per map kept in the compact list
constexpr int kSrcCap = 2 * kListCap;
constexpr int kBoxPix = 192;    // bounding boxes up to this area are evaluated pixel by pixel

struct DecodeParams {
  const float* maps;
  const float* maps_flip;
  const float* scal;
  const float* scal_flip;
  float* records;
  float* merged_out;
  int num_kpts;
  int count;  // batch * num_kpts maps
  int is_logits;
  int temp_is_pow2;
  float temperature, inv_temperature, normalize, err_div;
  int flip_idx[PP_MAX_KEYPOINTS];
  int radius[PP_MAX_KEYPOINTS];
  float taps[PP_MAX_KEYPOINTS][kTaps + 1];  // 1-D factor g of the OKS kernel, sum 1
};

__device__ __forceinline__ int reflect(int i, int n) {  // scipy 'reflect': d c b a | a b c d | d c b a
  return i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// One quadratic sub-pixel step on the convolved map (post_processing.py:384-430), fp32 like the reference.
__device__ __forceinline__ void subpixel(float c, float l, float r, float up, float dn, float& lx, float& ly) {
  const float dx = __fmul_rn(__fsub_rn(r, l), 0.5f), dy = __fmul_rn(__fsub_rn(dn, up), 0.5f);
  float dxx = __fsub_rn(__fadd_rn(r, l), __fmul_rn(2.f, c));
  float dyy = __fsub_rn(__fadd_rn(dn, up), __fmul_rn(2.f, c));
  if (dxx == 0.f) dxx = 1e-6f;
  if (dyy == 0.f) dyy = 1e-6f;
  lx = __fadd_rn(lx, __fdiv_rn(-dx, dxx));
  ly = __fadd_rn(ly, __fdiv_rn(-dy, dyy));
}

__device__ __forceinline__ void write_scalars(const DecodeParams& p, int b, int k, int kf, int j) {
  float s = 0.f;
  const int K = p.num_kpts;
  if (p.scal) {
    s = p.scal[(size_t)(b * 4 + j) * K + k];
    if (p.scal_flip) s = (s + p.scal_flip[(size_t)(b * 4 + j) * K + kf]) * 0.5f;
    if (j == 3) s = s / p.err_div;
  }
  p.records[(size_t)(b * K + k) * PP_RECORD_FLOATS + 3 + j] = s;
}

// =================================================================================================
// Dense path: the whole CTA on one map (any input).  Shared memory: three H x W planes + scratch.
// =================================================================================================
__device__ __forceinline__ void dense_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kDenseThreads) : "memory"); }

template <int N, typename T, typename Op>
__device__ __forceinline__ void block_allreduce(T (&v)[N], T (*scratch)[kDenseWarps][4], int& parity, Op op) {
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
  dense_sync();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T r = scratch[parity][0][i];
#pragma unroll
    for (int w = 1; w < kDenseWarps; ++w) r = op(r, scratch[parity][w][i]);
    v[i] = r;
  }
  parity ^= 1;
}

template <int H, int W>
__device__ __noinline__ void decode_dense(const DecodeParams& p, int item, float* sP, float* sH, float* sC,
                             float (*red_f)[kDenseWarps][4], int (*red_i)[kDenseWarps][4]) {
  constexpr int NPX = H * W;
  constexpr int NV4 = NPX / 4;
  constexpr int T = kDenseThreads;
  constexpr int V = NV4 / T;
  static_assert(NV4 % T == 0 && W % 4 == 0, "map must split into whole float4 per thread");
  int par_f = 0, par_i = 0;
  const int tid = threadIdx.x;
  const int K = p.num_kpts;
  const int b = item / K, k = item % K;
  const bool tta = p.maps_flip != nullptr;
  const int kf = tta ? p.flip_idx[k] : k;

  float z1[V][4], z2[V][4];
  {
    const float4* s1 = reinterpret_cast<const float4*>(p.maps + (size_t)(b * K + k) * NPX);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 t = ld_stream_f4(s1 + tid + j * T);
      z1[j][0] = t.x; z1[j][1] = t.y; z1[j][2] = t.z; z1[j][3] = t.w;
    }
    if (tta) {
      const float4* s2 = reinterpret_cast<const float4*>(p.maps_flip + (size_t)(b * K + kf) * NPX);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float4 t = ld_stream_f4(s2 + tid + j * T);
        z2[j][0] = t.x; z2[j][1] = t.y; z2[j][2] = t.z; z2[j][3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) z2[j][0] = z2[j][1] = z2[j][2] = z2[j][3] = 0.f;
    }
  }

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

  // flip-TTA merge into shared memory
#pragma unroll
  for (int j = 0; j < V; ++j)
    reinterpret_cast<float4*>(sP)[tid + j * T] = make_float4(z1[j][0], z1[j][1], z1[j][2], z1[j][3]);
  if (tta) {
    dense_sync();
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
  dense_sync();

  // support bounding box (any negative value -> treat the map as dense)
  int box[4] = {H, -1, W, -1};
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

  // separable OKS convolution over the dilated box, argmax
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
    dense_sync();
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
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red_f[par_f][warp][0] = best; red_i[par_i][warp][0] = best_i; }
    dense_sync();  // also orders the sC writes before thread 0 reads them
  }
  if (tid == 0) {
    for (int w = 0; w < kDenseWarps; ++w) {
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
    if (xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1)
      subpixel(cval(ys, xs), cval(ys, xs - 1), cval(ys, xs + 1), cval(ys - 1, xs), cval(ys + 1, xs), lx, ly);
    float* rec = p.records + (size_t)(b * K + k) * PP_RECORD_FLOATS;
    rec[0] = lx;
    rec[1] = ly;
    rec[2] = sP[best_i];
  }
  if (tid >= 32 && tid < 36) write_scalars(p, b, k, kf, tid - 32);
  dense_sync();  // the planes are reused by the next deferred map
}

// =================================================================================================
// Sparse path: one warp per map, the map streamed from global memory straight into registers.
// =================================================================================================
struct __align__(16) WarpList {
  float val[kSrcCap];            // candidate / source values
  unsigned short idx[kSrcCap];   // flat pixel index, later (y << 8 | x), of each candidate / source
  float taps[kTaps + 1];         // the map's OKS taps (zero-padded)
  float fac[kTaps + 1];          // weight * row factors of the source being scattered
};

__device__ __forceinline__ float warp_max_fast(float v) {  // CREDUX.MAX.F32 (sm_100a)
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// Streams one map (H x W fp32, 12 KB) through the warp's registers - 24 coalesced 128-bit loads per
// lane, only the maximum of every float4 is kept.  Returns the lane's mask of float4 that hold a
// candidate - logits: {z > max - T}, a superset of the sparsemax support; heatmaps: the positive
// pixels - the candidate threshold in `thr` and the maximum in `mx` (logits).  `bad` is set when the
// map has negative heatmap values (not a case for the sparse path).
template <int H, int W, bool LOGITS>
__device__ __forceinline__ unsigned warp_stream(const DecodeParams& p, const float* gmap, int lane, float& mx, float& thr,
                                                bool& bad) {
  constexpr int NV = H * W / 128;  // float4 per lane
  constexpr int CH = 6;            // loads in flight per lane and chunk
  static_assert(NV % CH == 0 && NV <= 32, "map must split into whole chunks of float4 per lane");
  const float4* g4 = reinterpret_cast<const float4*>(gmap) + lane;
  float m4[NV];
  float lo = 0.f;
  mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NV; c += CH) {
    float4 q[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) q[i] = __ldg(g4 + 32 * (c + i));
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      m4[c + i] = fmaxf(fmaxf(q[i].x, q[i].y), fmaxf(q[i].z, q[i].w));
      if (!LOGITS) lo = fminf(fminf(lo, q[i].x), fminf(fminf(q[i].y, q[i].z), q[i].w));
    }
  }
#pragma unroll
  for (int j = 0; j < NV; j += 2) mx = fmaxf(mx, fmaxf(m4[j], m4[j + 1]));
  thr = 0.f;
  bad = false;
  if (LOGITS) {
    mx = warp_max_fast(mx);
    // candidates: z / T > max / T - 1.  The raw-domain test is made slightly generous (any superset
    // of the support gives the same threshold); the exact scaled values are formed per candidate.
    thr = mx - p.temperature * 1.000001f - 1e-30f;
  } else {
    // negative values break the "maximum lies inside the support's bounding box" argument
    bad = __any_sync(0xffffffffu, lo < 0.f);
  }
  unsigned bits = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) bits |= (m4[j] > thr ? 1u : 0u) << j;
  return bits;
}

// Compacts the candidates of a streamed map into the list (idx, raw val) at list[off...]: in rounds,
// every lane re-reads its next float4 that holds a candidate (L1 / L2 hits); ballot compaction keeps
// the list order deterministic (round, component, lane).  Returns the number of entries, or -1 when
// there are too many for this path.
__device__ __forceinline__ int warp_compact(const float* gmap, unsigned bits, float thr, WarpList& wl, int off, int lane) {
  const float4* g4 = reinterpret_cast<const float4*>(gmap) + lane;
  const unsigned lt_mask = (1u << lane) - 1u;
  int n = 0;
  while (__any_sync(0xffffffffu, bits != 0u)) {
    const int j = __ffs(bits) - 1;
    float4 q = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (bits) q = __ldg(g4 + 32 * j);
    bits &= bits - 1u;
    const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool take = e[c] > thr;
      const unsigned bal = __ballot_sync(0xffffffffu, take);
      const int add = __popc(bal);
      if (n + add <= kListCap) {  // warp-uniform; on overflow only the count keeps growing
        if (take) {
          const int pos = off + n + __popc(bal & lt_mask);
          wl.val[pos] = e[c];
          wl.idx[pos] = (unsigned short)((lane + 32 * j) * 4 + c);
        }
      }
      n += add;
    }
    if (n > kListCap) break;  // does not fit: not sparse
  }
  __syncwarp();
  return n > kListCap ? -1 : n;
}

// List entries (pixel index, raw value) -> (y << 8 | x, heatmap value): sparsemax threshold by
// Michelot's fixed point on the list (logits), mirror of the flipped pass.
// Keeps the non-zero entries of list[off, off + n), order preserved; returns how many.
__device__ __forceinline__ int keep_support(WarpList& ws, int off, int n, int lane) {
  const unsigned lt_mask = (1u << lane) - 1u;
  int nnz = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const float v = i < n ? ws.val[off + i] : 0.f;
    const unsigned short yx = i < n ? ws.idx[off + i] : (unsigned short)0;
    const unsigned bal = __ballot_sync(0xffffffffu, v != 0.f);
    __syncwarp();  // the chunk is read before entries move down into it
    if (v != 0.f) {
      const int pos = off + nnz + __popc(bal & lt_mask);
      ws.val[pos] = v;
      ws.idx[pos] = yx;
    }
    nnz += __popc(bal);
  }
  __syncwarp();
  return nnz;
}

// Returns the length of the list after it has shrunk to its support (logits) / unchanged (heatmaps).
template <int H, int W, bool LOGITS>
__device__ __forceinline__ int warp_finish_list(const DecodeParams& p, WarpList& ws, int off, int n, int mirror, float mx,
                                                int lane) {
  for (int i = lane; i < n; i += 32) {
    const int px = ws.idx[off + i];
    const int y = px / W, x = px % W;
    ws.idx[off + i] = (unsigned short)((y << 8) | (mirror ? W - 1 - x : x));
  }
  if (LOGITS) {
    // tau <- (sum_{z > tau} z - 1) / #{z > tau} until the set stops shrinking = the sort / cumsum
    // threshold of Martins & Astudillo, Alg. 1 (on z - max z).
    const float mxs = p.temp_is_pow2 ? mx * p.inv_temperature : mx / p.temperature;
    auto scaled = [&](float e) { return (p.temp_is_pow2 ? e * p.inv_temperature : e / p.temperature) - mxs; };
    float tau = -1.f;
    if (n <= 32) {
      // the usual case, one candidate per lane: Martins & Astudillo Alg. 1 without the sort and without a
      // data-dependent iteration count - lane i counts k_i = #{z_j >= z_i} and S_i = sum of those z_j (= its position and
      // prefix sum in the sorted order; tied values share both), is in the support iff 1 + k_i z_i > S_i, and
      // tau = (S - 1) / k of the smallest supported value.  One pass over the list in shared memory.
      const float zi = lane < n ? scaled(ws.val[off + lane]) : -INFINITY;
      int k = 0;
      float sum = 0.f;
      for (int j = 0; j < n; ++j) {
        const float zj = scaled(ws.val[off + j]);
        if (zj >= zi) { ++k; sum += zj; }
      }
      const bool in = lane < n && (1.f + (float)k * zi > sum);
      const int kstar = __reduce_max_sync(0xffffffffu, in ? k : 0);  // >= 1: the maximum (z == 0) is always supported
      const unsigned who = __ballot_sync(0xffffffffu, in && k == kstar);
      const float sstar = __shfl_sync(0xffffffffu, sum, __ffs(who) - 1);
      tau = (sstar - 1.f) / (float)kstar;
      const float v = lane < n ? fminf(fmaxf(fmaxf(zi - tau, 0.f) * p.normalize, 0.f), 1.f) : 0.f;
      const unsigned short yx = lane < n ? ws.idx[off + lane] : (unsigned short)0;
      const unsigned bal = __ballot_sync(0xffffffffu, v != 0.f);
      __syncwarp();  // every lane has read the raw values and its index
      if (v != 0.f) {
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        ws.val[pos] = v;
        ws.idx[pos] = yx;
      }
      __syncwarp();
      return __popc(bal);
    } else {
      float zc[kListCap / 32];
#pragma unroll
      for (int i = 0; i < kListCap / 32; ++i) zc[i] = (lane + 32 * i < n) ? scaled(ws.val[off + lane + 32 * i]) : -INFINITY;
      float thr2 = -1.f;
      int prev = -1;
      for (int it = 0; it < kListCap + 2; ++it) {
        float sum = 0.f;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < kListCap / 32; ++i)
          if (zc[i] > thr2) { sum += zc[i]; ++cnt; }
        sum = warp_sum(sum);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        tau = (sum - 1.f) / (float)cnt;
        if (cnt == prev) break;
        prev = cnt;
        thr2 = fmaxf(thr2, tau);
      }
#pragma unroll
      for (int i = 0; i < kListCap / 32; ++i)
        if (lane + 32 * i < n) ws.val[off + lane + 32 * i] = fminf(fmaxf(fmaxf(zc[i] - tau, 0.f) * p.normalize, 0.f), 1.f);
      __syncwarp();
      return keep_support(ws, off, n, lane);
    }
  }
  __syncwarp();
  return n;
}

// First maximum of (value, flat index) pairs over the warp: CREDUX for the value, then the lowest index among the
// lanes that hold it.
__device__ __forceinline__ void warp_first_max(float& best, int& best_i) {
  const float m = warp_max_fast(best);
  best_i = __reduce_min_sync(0xffffffffu, best == m ? best_i : 0x7fffffff);
  best = m;
}

// The sparse decode of one map.  Returns false when the map has to go through the dense path.
template <int H, int W, bool LOGITS>
__device__ __forceinline__ bool decode_sparse(const DecodeParams& p, int item, WarpList& ws, int lane, float* tile) {
  constexpr int NPX = H * W;
  const int K = p.num_kpts;
  const int b = item / K, k = item % K;
  const bool tta = p.maps_flip != nullptr;
  const int kf = tta ? p.flip_idx[k] : k;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float* map1 = p.maps + (size_t)item * NPX;
  const float* map2 = tta ? p.maps_flip + (size_t)(b * K + kf) * NPX : nullptr;
  const int rad = p.radius[k];
  if (lane < kTaps + 1) ws.taps[lane] = p.taps[k][lane];  // read after the __syncwarp()s of the list stages

  float mx1, mx2 = 0.f, thr1, thr2 = 0.f;
  bool bad1, bad2 = false;
  const unsigned bits1 = warp_stream<H, W, LOGITS>(p, map1, lane, mx1, thr1, bad1);
  unsigned bits2 = 0;
  if (tta) bits2 = warp_stream<H, W, LOGITS>(p, map2, lane, mx2, thr2, bad2);
  if (bad1 || bad2) return false;

  const int n1 = warp_compact(map1, bits1, thr1, ws, 0, lane);
  if (n1 < 0) return false;
  int n2 = 0;
  if (tta) {
    n2 = warp_compact(map2, bits2, thr2, ws, n1, lane);
    if (n2 < 0) return false;
  }
  // both lists shrink to their supports first (a random-init head: ~15 candidates, ~4 supported); list 2 stays at n1
  const int m1 = warp_finish_list<H, W, LOGITS>(p, ws, 0, n1, 0, mx1, lane);
  int n = m1;
  if (tta) {
    const int m2 = warp_finish_list<H, W, LOGITS>(p, ws, n1, n2, 1, mx2, lane);
    // merged = (P + mirror(Pf)) * 0.5, pixel by pixel in fp32 exactly like the reference: an entry of
    // the first list absorbs the matching entry of the second; unmatched entries are halved on their own
    for (int i = lane; i < m1; i += 32) {
      const unsigned short yx = ws.idx[i];
      float other = 0.f;
      for (int j = n1; j < n1 + m2; ++j) other = ws.idx[j] == yx ? ws.val[j] : other;
      ws.val[i] = (ws.val[i] + other) * 0.5f;
    }
    __syncwarp();  // list-2 values are read above and rewritten below
    for (int i0 = 0; i0 < m2; i0 += 32) {  // unmatched entries of list 2 move up behind list 1 (matched ones as zeros)
      const int i = i0 + lane;
      float v = 0.f;
      unsigned short yx = 0;
      if (i < m2) {
        yx = ws.idx[n1 + i];
        bool dup = false;
        for (int j = 0; j < m1; ++j) dup |= ws.idx[j] == yx;
        v = dup ? 0.f : (0.f + ws.val[n1 + i]) * 0.5f;
      }
      __syncwarp();  // m1 + i <= n1 + i: a chunk never overwrites an entry a later chunk still has to read
      if (i < m2) {
        ws.val[m1 + i] = v;
        ws.idx[m1 + i] = yx;
      }
      __syncwarp();
    }
    n = m1 + m2;
  }

  // ---- keep the support only (non-zero entries), order preserved; bounding box ----
  int nnz = 0;
  int ymin = H, ymax = -1, xmin = W, xmax = -1;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const float v = i < n ? ws.val[i] : 0.f;
    const int yx = i < n ? ws.idx[i] : 0;
    const bool nz = v != 0.f;
    const unsigned bal = __ballot_sync(0xffffffffu, nz);
    __syncwarp();  // the chunk is read before entries move down into it
    if (nz) {
      const int pos = nnz + __popc(bal & lt_mask);
      ws.val[pos] = v;
      ws.idx[pos] = (unsigned short)yx;
      const int y = yx >> 8, x = yx & 255;
      ymin = min(ymin, y); ymax = max(ymax, y); xmin = min(xmin, x); xmax = max(xmax, x);
    }
    nnz += __popc(bal);
  }
  __syncwarp();
  ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
  xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
  const bool nonempty = nnz > 0;

  const float* gt = ws.taps + rad;
  auto g = [&](int d) { return gt[min(d, rad + 1)]; };  // g(d), 0 for d > rad (the table is zero-padded)
  int best_i = 0;
  float conf = 0.f, lx = 0.f, ly = 0.f;
  if (nonempty) {
    // C = sum_s w_s g(dy) g(dx) (+ the reflected images at -1 - s and 2H - 1 - s), evaluated directly
    auto eval_c = [&](int qy, int qx) -> float {
      float acc = 0.f;
      for (int s = 0; s < nnz; ++s) {
        const float w = ws.val[s];
        const int yx = ws.idx[s], sy = yx >> 8, sx = yx & 255;
        float fy = g(abs(qy - sy));
        float fx = g(abs(qx - sx));
        if (sy < rad || sy >= H - rad) {  // warp-uniform test
          fy += g(qy + 1 + sy);
          fy += g(2 * H - 1 - sy - qy);
        }
        if (sx < rad || sx >= W - rad) {
          fx += g(qx + 1 + sx);
          fx += g(2 * W - 1 - sx - qx);
        }
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(w, fy), fx));
      }
      return acc;
    };
    // The arg max of C lies inside the support's bounding box (the OKS kernel is non-negative and
    // decreasing in |d|), extended to the border where a reflected image can pull it outwards.
    const int sy0 = ymin <= rad - 1 ? 0 : ymin, sy1 = ymax >= H - rad ? H - 1 : ymax;
    const int sx0 = xmin <= rad - 1 ? 0 : xmin, sx1 = xmax >= W - rad ? W - 1 : xmax;
    const int bw = sx1 - sx0 + 1, area = (sy1 - sy0 + 1) * bw;
    float best = -INFINITY;
    best_i = 0x7fffffff;
    bool have_stencil = false;
    float st_c = 0.f, st_l = 0.f, st_r = 0.f, st_u = 0.f, st_d = 0.f;
    if (area <= kBoxPix) {
      // ---- compact support: every pixel of the box.  When the box dilated by one pixel (clamped to the map) fits the
      // warp, ONE evaluation pass yields the arg max and its four neighbours for the sub-pixel step ----
      const int dy0 = max(0, sy0 - 1), dy1 = min(H - 1, sy1 + 1), dx0 = max(0, sx0 - 1), dx1 = min(W - 1, sx1 + 1);
      const int dw = dx1 - dx0 + 1, darea = (dy1 - dy0 + 1) * dw;
      if (darea <= 32) {
        const int q = min(lane, darea - 1);
        const int qy = dy0 + q / dw, qx = dx0 + q % dw;
        const float c = eval_c(qy, qx);
        const bool inner = lane < darea && qy >= sy0 && qy <= sy1 && qx >= sx0 && qx <= sx1;  // the ring is not a candidate
        best = inner ? c : -INFINITY;
        best_i = inner ? qy * W + qx : 0x7fffffff;
        warp_first_max(best, best_i);
        if (best > 0.f) {
          const int ys = best_i / W, xs = best_i % W;
          if (xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1) {  // the four neighbours lie inside the dilated box
            const int li = (ys - dy0) * dw + (xs - dx0);
            st_c = __shfl_sync(0xffffffffu, c, li); st_l = __shfl_sync(0xffffffffu, c, li - 1); st_r = __shfl_sync(0xffffffffu, c, li + 1);
            st_u = __shfl_sync(0xffffffffu, c, li - dw); st_d = __shfl_sync(0xffffffffu, c, li + dw);
            have_stencil = true;
          }
        }
      } else {
        for (int q0 = 0; q0 < area; q0 += 32) {
          const int q = min(q0 + lane, area - 1);  // the duplicate of the last pixel never wins a tie
          const int qy = sy0 + q / bw, qx = sx0 + q % bw;
          const float c = eval_c(qy, qx);
          if (c > best) { best = c; best_i = qy * W + qx; }
        }
        warp_first_max(best, best_i);
      }
    } else {
      // ---- scattered support: every source adds w * g(dy) g(dx) (+ reflections) into the warp's 64 x 48
      // tile; the arg max is searched over the rows of the box ----
      // (measured: a pool of 3 tiles per 8 warps at twice the occupancy loses more to waiting than it gains)
      const int zy0 = max(0, ymin - rad - 1), zy1 = min(H - 1, ymax + rad + 1);
      float4* t4 = reinterpret_cast<float4*>(tile);
      for (int i = zy0 * (W / 4) + lane; i < (zy1 + 1) * (W / 4); i += 32) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      for (int s = 0; s < nnz; ++s) {
        const float w = ws.val[s];
        const int yx = ws.idx[s], sy = yx >> 8, sx = yx & 255;
        const int y0 = max(0, sy - rad), y1 = min(H - 1, sy + rad), x0 = max(0, sx - rad), x1 = min(W - 1, sx + rad);
        // 1-D factors over the window: the direct tap plus the taps of the two reflected images;
        // rows carry the weight, lanes keep their column factor in a register
        float fx = 0.f;
        if (lane <= 2 * rad) {
          const int yy = y0 + lane, xx = x0 + lane;
          if (yy <= y1) ws.fac[lane] = __fmul_rn(w, g(abs(yy - sy)) + g(yy + 1 + sy) + g(2 * H - 1 - sy - yy));
          if (xx <= x1) fx = g(abs(xx - sx)) + g(xx + 1 + sx) + g(2 * W - 1 - sx - xx);
        }
        __syncwarp();
        if (x0 + lane <= x1) {  // lane = window column; the window rows are independent read-modify-writes
          float* c = tile + y0 * W + x0 + lane;
          const int rows = y1 - y0 + 1;
          float cv[kTaps];
#pragma unroll
          for (int dy = 0; dy < kTaps; ++dy)
            if (dy < rows) cv[dy] = c[dy * W];
#pragma unroll
          for (int dy = 0; dy < kTaps; ++dy)
            if (dy < rows) c[dy * W] = __fadd_rn(cv[dy], __fmul_rn(ws.fac[dy], fx));
        }
        __syncwarp();
      }
      // first arg max over the rows of the search box
      for (int i = sy0 * (W / 4) + lane; i < (sy1 + 1) * (W / 4); i += 32) {
        const float4 c = t4[i];
        if (c.x > best) { best = c.x; best_i = 4 * i; }
        if (c.y > best) { best = c.y; best_i = 4 * i + 1; }
        if (c.z > best) { best = c.z; best_i = 4 * i + 2; }
        if (c.w > best) { best = c.w; best_i = 4 * i + 3; }
      }
      warp_first_max(best, best_i);
      if (best > 0.f) {
        const int ys = best_i / W, xs = best_i % W;
        if (xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1) {  // the stencil straight out of the tile: the same sums in the same order
          st_c = tile[best_i]; st_l = tile[best_i - 1]; st_r = tile[best_i + 1]; st_u = tile[best_i - W]; st_d = tile[best_i + W];
          have_stencil = true;
        }
      }
    }
    if (!(best > 0.f)) best_i = 0;  // weights underflowed: C == 0 everywhere, first maximum is pixel 0
    const int ys = best_i / W, xs = best_i % W;
    lx = (float)xs; ly = (float)ys;
    if (xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1) {
      if (!have_stencil) {
        // lanes 0..4: centre, left, right, up, down
        const int ddx = lane == 1 ? -1 : (lane == 2 ? 1 : 0), ddy = lane == 3 ? -1 : (lane == 4 ? 1 : 0);
        const float v = eval_c(ys + ddy, xs + ddx);
        st_c = __shfl_sync(0xffffffffu, v, 0); st_l = __shfl_sync(0xffffffffu, v, 1); st_r = __shfl_sync(0xffffffffu, v, 2);
        st_u = __shfl_sync(0xffffffffu, v, 3); st_d = __shfl_sync(0xffffffffu, v, 4);
      }
      subpixel(st_c, st_l, st_r, st_u, st_d, lx, ly);
    }
    // conf = merged heatmap at the integer peak (one entry per pixel after the merge)
    const int pyx = (ys << 8) | xs;
    float cv = 0.f;
    for (int i = lane; i < nnz; i += 32)
      if (ws.idx[i] == pyx) cv = ws.val[i];
    conf = warp_sum(cv);
  }

  if (p.merged_out) {
    float4* gout = reinterpret_cast<float4*>(p.merged_out + (size_t)item * NPX);
#pragma unroll 4
    for (int j = 0; j < NPX / 128; ++j) gout[lane + 32 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    for (int i = lane; i < nnz; i += 32)
      p.merged_out[(size_t)item * NPX + (ws.idx[i] >> 8) * W + (ws.idx[i] & 255)] = ws.val[i];
  }
  if (lane == 0) {
    float* rec = p.records + (size_t)item * PP_RECORD_FLOATS;
    rec[0] = lx;
    rec[1] = ly;
    rec[2] = conf;
  }
  if (lane >= 4 && lane < 8) write_scalars(p, b, k, kf, lane - 4);
  return true;
}

struct __align__(16) DecodeSmem {
  float planes[kDecWarps * 64 * 48];  // sparse path: one scatter tile per warp; dense path: P, row-convolved, C
  WarpList warp[kDecWarps];
  float red_f[2][kDenseWarps][4];
  int red_i[2][kDenseWarps][4];
  int queue[kDecWarps];
  int q_count;
};

// One warp per map, kDecWarps consecutive maps per CTA, kDecCtasPerSm CTAs per SM: 296 CTA slots on 148 SMs, so the
// 544 CTAs of a batch-256 call run in 1.84 waves; maps the sparse path declines are decoded afterwards by the CTA.
template <int H, int W, bool LOGITS>
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_kernel(const __grid_constant__ DecodeParams p) {
  constexpr int NPX = H * W;
  static_assert(NPX == 64 * 48, "planes are sized for 64 x 48 maps");
  extern __shared__ __align__(16) uint8_t dec_smem_raw[];
  DecodeSmem& sm = *reinterpret_cast<DecodeSmem*>(dec_smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  if (threadIdx.x == 0) sm.q_count = 0;
  __syncthreads();
  pdl_wait();  // the logits / scalars come from the previous kernels of the stream
  const int item = blockIdx.x * kDecWarps + warp;
  if (item < p.count) {
    if (!decode_sparse<H, W, LOGITS>(p, item, sm.warp[warp], lane, sm.planes + warp * NPX)) {
      if (lane == 0) sm.queue[atomicAdd(&sm.q_count, 1)] = item;
    }
  }
  __syncthreads();
  const int nq = sm.q_count;
  for (int q = 0; q < nq; ++q)
    decode_dense<H, W>(p, sm.queue[q], sm.planes, sm.planes + NPX, sm.planes + 2 * NPX, sm.red_f, sm.red_i);
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
  p.inv_temperature = 1.0f / cfg->temperature;
  {  // x / T == x * (1 / T) bit for bit when T is a power of two (the shipped 0.5)
    int ex = 0;
    p.temp_is_pow2 = cfg->input_is_logits && frexpf(cfg->temperature, &ex) == 0.5f;
  }
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

  const int64_t count = (int64_t)batch * cfg->num_keypoints;
  PP_REQUIRE(count < (1ll << 31), PP_ERR_INVALID, "pp_decode: batch too large");
  p.count = (int)count;
  static PerDeviceOnce attr_set;
  auto kern = cfg->input_is_logits ? decode_kernel<64, 48, true> : decode_kernel<64, 48, false>;
  if (attr_set.first()) {  // shared memory for kDecCtasPerSm CTAs per SM (about 110 KB each)
    PP_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel<64, 48, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeSmem)));
    PP_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel<64, 48, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeSmem)));
    PP_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel<64, 48, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PP_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel<64, 48, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  const int grid = (int)((count + kDecWarps - 1) / kDecWarps);
  PP_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kDecThreads), sizeof(DecodeSmem), (cudaStream_t)stream, p));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}
