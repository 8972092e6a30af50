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
// global write is the 28-byte record.  PERSISTENT CTAs, one per SM, each owning a contiguous range of maps:
//   * a PRODUCER warp (one elected thread) walks the range and feeds a kStages-deep ring of 12-KB shared-memory
//     stages with cp.async.bulk (1-D bulk copies, completion bytes on an mbarrier): the loads never wait for a
//     dependent chain,
//   * kConsumers CONSUMER warps claim maps in ring order (shared-memory counter), ONE WARP PER MAP: wait for the
//     map's stage(s), stream the stage through registers (24 conflict-free 128-bit shared-memory loads per lane,
//     only the maximum of each float4 is kept), compact the candidates into a short list and hand the stage back
//     to the producer at once - a stage is held for a few hundred instructions, the long list-only tail of the
//     decode (threshold, merge, convolution, arg max: ~1 500 latency-bound instructions) holds no stage, which is
//     what lets 31 warps per SM overlap their tails while 8 stages keep HBM busy,
//   * max (CREDUX) -> candidate set {z > max - T}: the few float4 that hold a candidate are re-read from the stage
//     and ballot-compacted into a short list (a trained head leaves 2-3 pixels) -> exact
//     sparsemax threshold on the list (all-pairs rank / prefix-sum form of Martins & Astudillo Alg. 1 for
//     <= 32 candidates, Michelot's fixed point beyond),
//   * the convolution is evaluated sparsely, C(q) = sum_s w_s g(dy) g(dx) (+ reflected images), only
//     where the arg max can be - inside the support's bounding box (the OKS kernel is non-negative and
//     decreasing, so the maximum of C cannot lie outside it): directly at the pixels of a compact box (the box
//     dilated by one pixel when that fits the warp, so that the sub-pixel stencil comes out of the same pass),
//     or - scattered supports - by scattering every source's window into a 64 x 48 tile (a small pool of tiles,
//     taken for the duration of the scatter + scan only) and scanning the rows of the box,
//   * all reductions are warp shuffles / redux: no block barrier on this path.
// Maps that are not sparse (flat random-init logits, arbitrary / negative heatmaps handed to the
// public codec API) are queued and decoded afterwards by the CTA with a dense separable
// convolution (decode_dense) - correct for any input, just not HBM-bound.
#include "common.cuh"
#include "ptx.cuh"

#include <math.h>

namespace pp {

constexpr int kMaxRadius = 9;  // ceil(3 * 3.0): the variance is clipped to <= 3.0
constexpr int kTaps = 2 * kMaxRadius + 1;
constexpr int kConsumers = 20;  // consumer warps (one map each at a time)
constexpr int kStages = 12;     // ring depth: 12-KB stages, held only while a map is streamed and its candidates compacted
constexpr int kTiles = 4;       // 12-KB scatter tiles for maps with scattered supports (acquired after the stage is released)
constexpr int kFullBars = 128;  // "landed" barriers, indexed by copy number: >= kConsumers * 2 + kStages (see the kernel)
constexpr int kDecThreads = 32 * (kConsumers + 1);  // + the producer warp
constexpr int kDenseWarps = 8;   // the dense path uses 8 warps (3 float4 of the map per thread)
constexpr int kDenseThreads = 32 * kDenseWarps;
constexpr int kListCap = 96;    // candidates per map kept in the compact list
constexpr int kSrcCap = 2 * kListCap;
constexpr int kBoxPix = 192;    // a bounding box up to this area is evaluated pixel by pixel
constexpr int kHopPix = 384;    // ... and so are the one-hop boxes of a scattered support up to this many pixels in total
constexpr int kQueueCap = 512;  // maps per CTA (bounds the dense-path queue)

struct DecodeParams {
  const float* maps;
  const float* maps_flip;
  const float* scal;
  const float* scal_flip;
  float* records;
  float* merged_out;
  int num_kpts;
  int count;          // batch * num_kpts maps
  int items_per_cta;  // contiguous maps per CTA (<= kQueueCap)
  int probe;          // profiling only (PP_DECODE_PROBE): 1 = consumers only release the stages, 2 = stop after the candidate lists
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
// Sparse path: one warp per map, the map staged in shared memory by the producer's bulk copy.
// =================================================================================================
struct __align__(16) WarpList {
  float val[kSrcCap];            // candidate / source values
  unsigned short idx[kSrcCap];   // flat pixel index, later (y << 8 | x), of each candidate / source
  float taps[kTaps + 1];         // the map's OKS taps (zero-padded)
  float fac[kTaps + 1];          // weight * row factors of the source being scattered
  unsigned box[32];              // search boxes (y0 | y1 << 8 | x0 << 16 | x1 << 24)
  int pre[33];                   // pixels before box i
};

__device__ __forceinline__ float warp_max_fast(float v) {  // CREDUX.MAX.F32 (sm_100a)
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// 1-D bulk copy global -> shared (async proxy), completion bytes on an mbarrier.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
// Generic-proxy writes to a stage (the scatter tile) are ordered before the producer's next bulk copy into it.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- code size ----------------------------------------------------------------------------------------------
// Every map walks this code ONCE, 20 warps per SM are at different places in it, and a straight-line pass over
// thousands of instructions runs at the speed of the instruction cache refill (measured: 14 000 cycles per map with
// the loops unrolled and the helpers inlined at every call site, most of it "no instruction" stalls).  Hence: rolled
// loops wherever four loads in flight are enough, the per-map stages as __noinline__ functions shared by both passes
// of a TTA call, and the rare paths (more than 32 candidates, scattered supports, the scatter tile) out of line.

// Scans one staged map (H x W fp32, 12 KB): maximum (CREDUX), then the candidates - logits: {z > max - T}, a
// superset of the sparsemax support; heatmaps: the positive pixels - ballot-compacted into the list at wl[off...] as
// (y << 8 | x, raw value), x mirrored for the flipped pass.  Deterministic order (float4 index, component, lane).
// Returns the number of candidates, or -1 when the map is not a case for the sparse path (too many candidates,
// negative heatmap values).
template <int H, int W, bool LOGITS>
__device__ __noinline__ int scan_map(const DecodeParams& p, const float* smap, WarpList& wl, int off, int mirror, float& mx_out) {
  constexpr int NV = H * W / 128;  // float4 per lane
  static_assert(NV <= 32 && W % 4 == 0, "one candidate bit per float4 of the lane; a float4 never straddles two rows");
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float4* s4 = reinterpret_cast<const float4*>(smap) + lane;
  // the one place where full unrolling pays: 24 independent 128-bit loads in flight, only the maximum of each kept
  float m4[NV];
  float lo = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 q = s4[32 * j];
    m4[j] = fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w));
    if (!LOGITS) lo = fminf(fminf(lo, q.x), fminf(fminf(q.y, q.z), q.w));
  }
  float mx = -INFINITY, thr = 0.f;
#pragma unroll
  for (int j = 0; j < NV; j += 2) mx = fmaxf(mx, fmaxf(m4[j], m4[j + 1]));
  if (LOGITS) {
    mx = warp_max_fast(mx);
    // candidates: z / T > max / T - 1.  The raw-domain test is made slightly generous (any superset
    // of the support gives the same threshold); the exact scaled values are formed per candidate.
    thr = mx - p.temperature * 1.000001f - 1e-30f;
  } else if (__any_sync(0xffffffffu, lo < 0.f)) {
    return -1;  // negative values break the "maximum lies inside the support's bounding box" argument
  }
  mx_out = mx;
  unsigned bits = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) bits |= (m4[j] > thr ? 1u : 0u) << j;
  // in rounds, every lane re-reads its next float4 that holds a candidate; ballot compaction keeps the list order
  // deterministic (round, component, lane)
  int n = 0;
#pragma unroll 1
  while (__any_sync(0xffffffffu, bits != 0u)) {
    const int j = __ffs(bits) - 1;
    float4 q = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (bits) q = s4[32 * j];
    bits &= bits - 1u;
    const int pix = (lane + 32 * j) * 4, y = pix / W, x = pix % W;
    const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool take = e[c] > thr;
      const unsigned bal = __ballot_sync(0xffffffffu, take);
      const int add = __popc(bal);
      if (take && n + add <= kListCap) {  // on overflow only the count keeps growing
        const int pos = off + n + __popc(bal & lt_mask);
        wl.val[pos] = e[c];
        wl.idx[pos] = (unsigned short)((y << 8) | (mirror ? W - 1 - (x + c) : x + c));
      }
      n += add;
    }
    if (n > kListCap) return -1;  // does not fit: not sparse
  }
  __syncwarp();
  return n;
}

// More than 32 candidates (rare): tau <- (sum_{z > tau} z - 1) / #{z > tau} until the set stops shrinking (Michelot)
// = the sort / cumsum threshold of Martins & Astudillo, Alg. 1 (on z - max z).
// Keeps the non-zero entries of list[off, off + n), order preserved; returns how many.
__device__ __forceinline__ int keep_support(WarpList& ws, int off, int n) {
  const int lane = threadIdx.x & 31;
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

__device__ __noinline__ void sparsemax_list_big(const DecodeParams& p, WarpList& ws, int off, int n, float mx) {
  const int lane = threadIdx.x & 31;
  const float mxs = p.temp_is_pow2 ? mx * p.inv_temperature : mx / p.temperature;
  auto scaled = [&](float e) { return (p.temp_is_pow2 ? e * p.inv_temperature : e / p.temperature) - mxs; };
  float thr2 = -1.f, tau = -1.f;
  int prev = -1;
  for (int it = 0; it < kListCap + 2; ++it) {
    float sum = 0.f;
    int cnt = 0;
    for (int i = lane; i < n; i += 32) {
      const float z = scaled(ws.val[off + i]);
      if (z > thr2) { sum += z; ++cnt; }
    }
    sum = warp_sum(sum);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    tau = (sum - 1.f) / (float)cnt;  // cnt >= 1: the maximum (z == 0) is always a candidate
    if (cnt == prev) break;
    prev = cnt;
    thr2 = fmaxf(thr2, tau);
  }
  for (int i = lane; i < n; i += 32)
    ws.val[off + i] = fminf(fmaxf(fmaxf(scaled(ws.val[off + i]) - tau, 0.f) * p.normalize, 0.f), 1.f);
  __syncwarp();
}

// Raw candidate values -> heatmap values: exact sparsemax threshold on the list; the list shrinks to its support
// (returns the new length).
__device__ __noinline__ int sparsemax_list(const DecodeParams& p, WarpList& ws, int off, int n, float mx) {
  if (n > 32) {
    sparsemax_list_big(p, ws, off, n, mx);
    return keep_support(ws, off, n);
  }
  // the usual case, one candidate per lane: Martins & Astudillo Alg. 1 without the sort - lane i counts
  // k_i = #{z_j >= z_i} and S_i = sum of those z_j (= its position and prefix sum in the sorted order; tied
  // values share both), is in the support iff 1 + k_i z_i > S_i, and tau = (S - 1) / k of the smallest
  // supported value.  No data-dependent iteration count: one pass over the list in shared memory.
  const int lane = threadIdx.x & 31;
  const float mxs = p.temp_is_pow2 ? mx * p.inv_temperature : mx / p.temperature;
  auto scaled = [&](float e) { return (p.temp_is_pow2 ? e * p.inv_temperature : e / p.temperature) - mxs; };
  const float zi = lane < n ? scaled(ws.val[off + lane]) : -INFINITY;
  int k = 0;
  float sum = 0.f;
#pragma unroll 2
  for (int j = 0; j < n; ++j) {
    const float zj = scaled(ws.val[off + j]);
    if (zj >= zi) { ++k; sum += zj; }
  }
  const bool in = lane < n && (1.f + (float)k * zi > sum);
  const int kstar = __reduce_max_sync(0xffffffffu, in ? k : 0);  // >= 1: the maximum (z == 0) is always supported
  const unsigned who = __ballot_sync(0xffffffffu, in && k == kstar);
  const float sstar = __shfl_sync(0xffffffffu, sum, __ffs(who) - 1);
  const float tau = (sstar - 1.f) / (float)kstar;
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
}

// First maximum of (value, flat index) pairs over the warp: CREDUX for the value, then the lowest index among the
// lanes that hold it.
__device__ __forceinline__ void warp_first_max(float& best, int& best_i) {
  const float m = warp_max_fast(best);
  best_i = __reduce_min_sync(0xffffffffu, best == m ? best_i : 0x7fffffff);
  best = m;
}

#ifdef PP_DECODE_TIMING
// Debug build only: per-phase cycle totals of the sparse decode (lane 0 of every consumer warp), read back through
// pp_debug_decode_timing().  [0] items, [1] wait for the stage(s), [2] scan (map), [3] scan (flipped map),
// [4] sparsemax + merge + support, [5] search set-up, [6] evaluation / tile, [7] everything after the scans.
__device__ unsigned long long g_dec_timing[8];
#define PP_T(var) const long long var = clock64()
#define PP_TADD(slot, a, b) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_dec_timing[slot], (unsigned long long)((b) - (a))); } while (0)
#else
#define PP_T(var)
#define PP_TADD(slot, a, b)
#endif

// Scatter tiles: a bit mask of free tiles in shared memory; a warp takes one for its scatter + scan only.
__device__ __forceinline__ int tile_acquire(unsigned* free_mask, int lane) {
  int t = 0;
  if (lane == 0) {
    for (;;) {
      const unsigned m = *reinterpret_cast<volatile unsigned*>(free_mask);
      if (m == 0u) { __nanosleep(100); continue; }
      t = __ffs(m) - 1;
      if (atomicAnd(free_mask, ~(1u << t)) & (1u << t)) break;
    }
  }
  return __shfl_sync(0xffffffffu, t, 0);
}
__device__ __forceinline__ void tile_release(unsigned* free_mask, int t, int lane) {
  __syncwarp();
  if (lane == 0) atomicOr(free_mask, 1u << t);
}

struct Support {  // the merged map's non-zero pixels: ws.val / ws.idx [0, nnz), bounding box, the keypoint's OKS radius
  int nnz, ymin, ymax, xmin, xmax, rad;
};

__device__ __forceinline__ float oks_tap(const WarpList& ws, int rad, int d) {  // g(d), 0 for d > rad (zero-padded table)
  return ws.taps[rad + min(d, rad + 1)];
}

// The box of the arg max is extended to the border where a reflected image can pull the maximum outwards.
template <int H, int W>
__device__ __forceinline__ void extend_box(int rad, int& y0, int& y1, int& x0, int& x1) {
  if (y0 <= rad - 1) y0 = 0;
  if (y1 >= H - rad) y1 = H - 1;
  if (x0 <= rad - 1) x0 = 0;
  if (x1 >= W - rad) x1 = W - 1;
}

// Scattered support: for every source the box of the sources whose windows can overlap its own (Chebyshev
// distance <= 2 rad).  Any pixel q sees only sources that are mutual neighbours (all within rad of q); moving q
// into their box does not decrease any of their terms and the others are >= 0 - so the maximum over the union of
// these one-hop boxes is the global one, no connected components needed.  An isolated source contributes a single
// pixel.  Boxes contained in another source's box are dropped.  Fills ws.box / ws.pre, returns the pixel total.
template <int H, int W>
__device__ __noinline__ int one_hop_boxes(WarpList& ws, const Support& sp) {
  const int lane = threadIdx.x & 31, nnz = sp.nnz, rad = sp.rad;
  int y0 = 0, y1 = 0, x0 = 0, x1 = 0;
  if (lane < nnz) {
    const int yx = ws.idx[lane];
    const int sy = yx >> 8, sx = yx & 255;
    y0 = y1 = sy; x0 = x1 = sx;
    for (int j = 0; j < nnz; ++j) {
      const int o = ws.idx[j], oy = o >> 8, ox = o & 255;
      if (abs(oy - sy) <= 2 * rad && abs(ox - sx) <= 2 * rad) {
        y0 = min(y0, oy); y1 = max(y1, oy); x0 = min(x0, ox); x1 = max(x1, ox);
      }
    }
    extend_box<H, W>(rad, y0, y1, x0, x1);
    ws.box[lane] = (unsigned)y0 | ((unsigned)y1 << 8) | ((unsigned)x0 << 16) | ((unsigned)x1 << 24);
  }
  __syncwarp();
  int my_area = 0;
  if (lane < nnz) {
    bool drop = false;
    for (int j = 0; j < nnz; ++j) {
      const unsigned o = ws.box[j];
      const int oy0 = o & 255, oy1 = (o >> 8) & 255, ox0 = (o >> 16) & 255, ox1 = o >> 24;
      const bool inside = oy0 <= y0 && oy1 >= y1 && ox0 <= x0 && ox1 >= x1;
      const bool same = oy0 == y0 && oy1 == y1 && ox0 == x0 && ox1 == x1;
      drop |= j != lane && inside && (!same || j < lane);
    }
    my_area = drop ? 0 : (y1 - y0 + 1) * (x1 - x0 + 1);
  }
  int incl = my_area;  // inclusive prefix sum over the lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane < nnz) ws.pre[lane + 1] = incl;
  if (lane == 0) ws.pre[0] = 0;
  __syncwarp();
  return __shfl_sync(0xffffffffu, incl, 31);
}

// Scatter tile (big overlapping windows): every source adds w * g(dy) g(dx) (+ reflections) into a 64 x 48 tile
// from the pool; first arg max over the rows of the search box, the sub-pixel stencil straight out of the tile.
// Returns true when the stencil is valid.
template <int H, int W>
__device__ __noinline__ bool tile_search(WarpList& ws, const Support& sp, int sy0, int sy1, float* tiles, unsigned* tile_mask,
                                         float& best, int& best_i, float (&st)[5]) {
  const int lane = threadIdx.x & 31, rad = sp.rad;
  const int tile_id = tile_acquire(tile_mask, lane);
  float* tile = tiles + (size_t)tile_id * (H * W);
  const int zy0 = max(0, sp.ymin - rad - 1), zy1 = min(H - 1, sp.ymax + rad + 1);
  float4* t4 = reinterpret_cast<float4*>(tile);
  for (int i = zy0 * (W / 4) + lane; i < (zy1 + 1) * (W / 4); i += 32) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
#pragma unroll 1
  for (int s = 0; s < sp.nnz; ++s) {
    const float w = ws.val[s];
    const int yx = ws.idx[s], sy = yx >> 8, sx = yx & 255;
    const int y0 = max(0, sy - rad), y1 = min(H - 1, sy + rad), x0 = max(0, sx - rad), x1 = min(W - 1, sx + rad);
    // 1-D factors over the window: the direct tap plus the taps of the two reflected images (only where a
    // reflected image is within the radius: warp-uniform tests); rows carry the weight, lanes keep their
    // column factor in a register
    float fx = 0.f;
    if (lane <= 2 * rad) {
      const int yy = y0 + lane, xx = x0 + lane;
      if (yy <= y1) {
        float fy = oks_tap(ws, rad, abs(yy - sy));
        if (sy < rad || sy >= H - rad) fy = fy + oks_tap(ws, rad, yy + 1 + sy) + oks_tap(ws, rad, 2 * H - 1 - sy - yy);
        ws.fac[lane] = __fmul_rn(w, fy);
      }
      if (xx <= x1) {
        fx = oks_tap(ws, rad, abs(xx - sx));
        if (sx < rad || sx >= W - rad) fx = fx + oks_tap(ws, rad, xx + 1 + sx) + oks_tap(ws, rad, 2 * W - 1 - sx - xx);
      }
    }
    __syncwarp();
    if (x0 + lane <= x1) {  // lane = window column; the window rows are independent read-modify-writes
      float* c = tile + y0 * W + x0 + lane;
      const int rows = y1 - y0 + 1;
      int dy = 0;
#pragma unroll 1
      for (; dy + 4 <= rows; dy += 4) {  // four rows in flight
        const float c0 = c[dy * W], c1 = c[(dy + 1) * W], c2 = c[(dy + 2) * W], c3 = c[(dy + 3) * W];
        const float f0 = ws.fac[dy], f1 = ws.fac[dy + 1], f2 = ws.fac[dy + 2], f3 = ws.fac[dy + 3];
        c[dy * W] = __fadd_rn(c0, __fmul_rn(f0, fx));
        c[(dy + 1) * W] = __fadd_rn(c1, __fmul_rn(f1, fx));
        c[(dy + 2) * W] = __fadd_rn(c2, __fmul_rn(f2, fx));
        c[(dy + 3) * W] = __fadd_rn(c3, __fmul_rn(f3, fx));
      }
#pragma unroll 1
      for (; dy < rows; ++dy) c[dy * W] = __fadd_rn(c[dy * W], __fmul_rn(ws.fac[dy], fx));
    }
    __syncwarp();
  }
  // the maximum of each float4 first, the component only on a new best
#pragma unroll 2
  for (int i = sy0 * (W / 4) + lane; i < (sy1 + 1) * (W / 4); i += 32) {
    const float4 c = t4[i];
    const float m = fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w));
    if (m > best) {
      best = m;
      best_i = 4 * i + (c.x == m ? 0 : (c.y == m ? 1 : (c.z == m ? 2 : 3)));
    }
  }
  warp_first_max(best, best_i);
  bool have = false;
  if (best > 0.f) {
    const int ys = best_i / W, xs = best_i % W;
    if (xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1) {  // same sums, same order as any other evaluation of C
      st[0] = tile[best_i]; st[1] = tile[best_i - 1]; st[2] = tile[best_i + 1]; st[3] = tile[best_i - W]; st[4] = tile[best_i + W];
      have = true;
    }
  }
  tile_release(tile_mask, tile_id, lane);
  return have;
}

// The sparse decode of one map whose stage(s) have landed.  `release(which)` hands stage 0 (the map) / 1 (the
// flipped pass's map) back to the producer as soon as its candidates are in the list; every stage is released
// exactly once on every path.  Returns false when the map has to go through the dense path.
template <int H, int W, bool LOGITS, typename Release>
__device__ __forceinline__ bool decode_sparse(const DecodeParams& p, int item, WarpList& ws, int lane, const float* stage1,
                                              const float* stage2, float* tiles, unsigned* tile_mask, Release&& release) {
  constexpr int NPX = H * W;
  const int K = p.num_kpts;
  const int b = item / K, k = item % K;
  const bool tta = p.maps_flip != nullptr;
  const int kf = tta ? p.flip_idx[k] : k;
  const unsigned lt_mask = (1u << lane) - 1u;
  Support sp;
  sp.rad = p.radius[k];
  const int rad = sp.rad;
  if (lane < kTaps + 1) ws.taps[lane] = p.taps[k][lane];  // read after the __syncwarp()s of the list stages

  float mx1 = 0.f, mx2 = 0.f;
  PP_T(t_a0);
  const int n1 = scan_map<H, W, LOGITS>(p, stage1, ws, 0, 0, mx1);
  release(0);  // the map's candidates are in the list (or the map goes to the dense path)
  PP_T(t_a1);
  PP_TADD(2, t_a0, t_a1);
  int n2 = 0;
  if (tta) {
    if (n1 >= 0) n2 = scan_map<H, W, LOGITS>(p, stage2, ws, n1, 1, mx2);
    release(1);
  }
  PP_T(t_a2);
  PP_TADD(3, t_a1, t_a2);
  if (n1 < 0 || n2 < 0) return false;
  if (p.probe == 2) return true;
  // both lists shrink to their supports first (a random-init head: ~15 candidates, ~4 supported), list 2 stays at n1
  const int m1 = LOGITS ? sparsemax_list(p, ws, 0, n1, mx1) : n1;
  int n = m1;
  if (tta) {
    const int m2 = LOGITS ? sparsemax_list(p, ws, n1, n2, mx2) : n2;
    // merged = (P + mirror(Pf)) * 0.5, pixel by pixel in fp32 exactly like the reference: an entry of
    // the first list absorbs the matching entry of the second; unmatched entries are halved on their own
    for (int i = lane; i < m1; i += 32) {
      const unsigned short yx = ws.idx[i];
      float other = 0.f;
      for (int j = n1; j < n1 + m2; ++j) other = ws.idx[j] == yx ? ws.val[j] : other;
      ws.val[i] = (ws.val[i] + other) * 0.5f;
    }
    __syncwarp();  // list-2 values are read above and rewritten below
    for (int i0 = 0; i0 < m2; i0 += 32) {  // unmatched entries of list 2 move up behind list 1
      const int i = i0 + lane;
      float v = 0.f;
      unsigned short yx = 0;
      if (i < m2) {
        yx = ws.idx[n1 + i];
        bool dup = false;
        for (int j = 0; j < m1; ++j) dup |= ws.idx[j] == yx;
        v = dup ? 0.f : (0.f + ws.val[n1 + i]) * 0.5f;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, i < m2);
      __syncwarp();
      if (i < m2) {
        ws.val[m1 + i] = v;  // m1 + i <= n1 + i: never ahead of an unread entry of this or a later chunk
        ws.idx[m1 + i] = yx;
      }
      (void)bal;
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
  sp.nnz = nnz; sp.ymin = ymin; sp.ymax = ymax; sp.xmin = xmin; sp.xmax = xmax;
  const bool nonempty = nnz > 0;
  PP_T(t_b0);
  PP_TADD(4, t_a2, t_b0);

  int best_i = 0;
  float conf = 0.f, lx = 0.f, ly = 0.f;
  if (nonempty) {
    // The arg max of C lies inside the support's bounding box (the OKS kernel is non-negative and
    // decreasing in |d|), extended to the border where a reflected image can pull it outwards.
    int sy0 = ymin, sy1 = ymax, sx0 = xmin, sx1 = xmax;
    extend_box<H, W>(rad, sy0, sy1, sx0, sx1);
    const int area = (sy1 - sy0 + 1) * (sx1 - sx0 + 1);
    float best = -INFINITY;
    best_i = 0x7fffffff;
    bool have_stencil = false;
    float st[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // C at the peak, left, right, up, down

    // ---- where to look: a list of boxes (packed y0 | y1 << 8 | x0 << 16 | x1 << 24 in ws.box, running pixel counts in
    // ws.pre) that is guaranteed to hold the arg max: compact support - the bounding box itself (dilated by one pixel
    // when that fits the warp, so that one evaluation pass also yields the four neighbours of the sub-pixel step);
    // scattered support - one_hop_boxes(); too many pixels that way - the scatter tile.
    int total = 0, nboxes = 1;
    bool fused = false;  // single dilated box of <= 32 pixels
    bool use_tile = false;
    if (area <= kBoxPix) {
      int dy0 = max(0, sy0 - 1), dy1 = min(H - 1, sy1 + 1), dx0 = max(0, sx0 - 1), dx1 = min(W - 1, sx1 + 1);
      fused = (dy1 - dy0 + 1) * (dx1 - dx0 + 1) <= 32;
      if (!fused) { dy0 = sy0; dy1 = sy1; dx0 = sx0; dx1 = sx1; }
      total = (dy1 - dy0 + 1) * (dx1 - dx0 + 1);
      if (lane == 0) {
        ws.box[0] = (unsigned)dy0 | ((unsigned)dy1 << 8) | ((unsigned)dx0 << 16) | ((unsigned)dx1 << 24);
        ws.pre[0] = 0;
        ws.pre[1] = total;
      }
      __syncwarp();
    } else if (nnz <= 32) {
      total = one_hop_boxes<H, W>(ws, sp);
      nboxes = nnz;
      use_tile = total > (p.probe >= 100 ? p.probe - 100 : kHopPix);
    } else {
      use_tile = true;
    }
    PP_T(t_b1);
    PP_TADD(5, t_b0, t_b1);

    if (!use_tile) {
      // one loop evaluates the pixels of the box list in rounds of 32 and, in a last extra round, the five pixels of
      // the sub-pixel stencil (unless the fused case already has them): the evaluation of
      // C = sum_s w_s g(dy) g(dx) (+ the reflected images at -1 - s and 2H - 1 - s) exists once in the code
      const int rounds = (total + 31) >> 5;
      float c_fused = 0.f;
      int fdy0 = 0, fdx0 = 0, fdw = 1;
#pragma unroll 1
      for (int r = 0; r <= rounds; ++r) {
        int qy, qx;
        bool valid = false;
        if (r < rounds) {
          const int gi = min(r * 32 + lane, total - 1);  // a duplicate of the last pixel never wins a tie
          int bsel = 0;
          while (bsel + 1 < nboxes && gi >= ws.pre[bsel + 1]) ++bsel;
          const unsigned bb = ws.box[bsel];
          const int y0 = bb & 255, x0 = (bb >> 16) & 255, bw = (int)(bb >> 24) - x0 + 1;
          const int li = gi - ws.pre[bsel];
          const int row = (int)(((float)li + 0.5f) * __frcp_rn((float)bw));  // exact: li < 4096, bw <= 48
          qy = y0 + row;
          qx = x0 + li - row * bw;
          valid = !fused || (qy >= sy0 && qy <= sy1 && qx >= sx0 && qx <= sx1);  // the dilation ring is not a candidate
          if (fused) { fdy0 = y0; fdx0 = x0; fdw = bw; }
        } else {
          if (have_stencil || !(best > 0.f)) break;
          const int ys = best_i / W, xs = best_i % W;
          if (!(xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1)) break;
          // lanes 0..4: centre, left, right, up, down
          qx = xs + (lane == 1 ? -1 : (lane == 2 ? 1 : 0));
          qy = ys + (lane == 3 ? -1 : (lane == 4 ? 1 : 0));
        }
        float c = 0.f;
#pragma unroll 1
        for (int s = 0; s < nnz; ++s) {
          const float w = ws.val[s];
          const int yx = ws.idx[s], sy = yx >> 8, sx = yx & 255;
          float fy = oks_tap(ws, rad, abs(qy - sy));
          float fx = oks_tap(ws, rad, abs(qx - sx));
          if (sy < rad || sy >= H - rad) {  // warp-uniform test
            fy += oks_tap(ws, rad, qy + 1 + sy);
            fy += oks_tap(ws, rad, 2 * H - 1 - sy - qy);
          }
          if (sx < rad || sx >= W - rad) {
            fx += oks_tap(ws, rad, qx + 1 + sx);
            fx += oks_tap(ws, rad, 2 * W - 1 - sx - qx);
          }
          c = __fadd_rn(c, __fmul_rn(__fmul_rn(w, fy), fx));
        }
        if (r < rounds) {
          if (valid && c > best) { best = c; best_i = qy * W + qx; }
          if (fused) c_fused = c;
          if (r == rounds - 1) {
            warp_first_max(best, best_i);
            if (fused && best > 0.f) {
              const int ys = best_i / W, xs = best_i % W;
              if (xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1) {  // the four neighbours lie inside the dilated box
                const int li = (ys - fdy0) * fdw + (xs - fdx0);
                st[0] = __shfl_sync(0xffffffffu, c_fused, li); st[1] = __shfl_sync(0xffffffffu, c_fused, li - 1);
                st[2] = __shfl_sync(0xffffffffu, c_fused, li + 1); st[3] = __shfl_sync(0xffffffffu, c_fused, li - fdw);
                st[4] = __shfl_sync(0xffffffffu, c_fused, li + fdw);
                have_stencil = true;
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 5; ++i) st[i] = __shfl_sync(0xffffffffu, c, i);
          have_stencil = true;
        }
      }
    } else {
      have_stencil = tile_search<H, W>(ws, sp, sy0, sy1, tiles, tile_mask, best, best_i, st);
    }
    PP_T(t_b2);
    PP_TADD(6, t_b1, t_b2);
    if (!(best > 0.f)) best_i = 0;  // weights underflowed: C == 0 everywhere, first maximum is pixel 0
    const int ys = best_i / W, xs = best_i % W;
    lx = (float)xs; ly = (float)ys;
    if (have_stencil && xs > 0 && xs < W - 1 && ys > 0 && ys < H - 1) subpixel(st[0], st[1], st[2], st[3], st[4], lx, ly);
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
  PP_T(t_c);
  PP_TADD(7, t_a2, t_c);
  PP_TADD(0, 0, 1);
  return true;
}

struct __align__(128) DecodeSmem {
  float stage[kStages][64 * 48];  // the ring; dense path (after the ring has drained): P, row-convolved, C = stages 0..2
  float tile[kTiles][64 * 48];    // scatter tiles
  WarpList warp[kConsumers];
  uint64_t full[kFullBars], empty[kStages];
  float red_f[2][kDenseWarps][4];
  int red_i[2][kDenseWarps][4];
  unsigned short queue[kQueueCap];  // maps of this CTA (relative index) the sparse path declined
  int q_count;
  int next_item;
  unsigned tile_mask;  // free scatter tiles
};
static_assert(sizeof(DecodeSmem) <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
static_assert(kFullBars >= 2 * kConsumers + kStages, "see the comment on the landed barriers");

// Persistent CTA: warp kConsumers = producer (bulk copies into the ring), warps 0 .. kConsumers-1 = consumers
// (one map each at a time, claimed in ring order); maps the sparse path declines are decoded afterwards by
// warps 0 .. kDenseWarps-1 together.
//
// Barriers.  empty[s]: the producer is its only waiter and walks the ring in order - plain phase parity.  The "landed"
// barriers are indexed by COPY NUMBER (n mod kFullBars), not by stage: maps are claimed in ring order but finish out
// of order, so a consumer may start waiting for copy n while an earlier use of the same stage has not even been issued,
// and a parity wait can only tell a barrier's current phase from the one before it.  Copy n - kFullBars must have
// landed before anybody waits for copy n: otherwise no copy >= n - kFullBars + kStages has been issued (the ring is
// in order), and each of the (kFullBars - kStages) / 2 >= kConsumers maps claimed in between would be pinning one
// consumer warp in its wait - more warps than there are.
template <int H, int W, bool LOGITS>
__global__ void __launch_bounds__(kDecThreads, 1) decode_kernel(const __grid_constant__ DecodeParams p) {
  constexpr int NPX = H * W;
  constexpr uint32_t MAP_BYTES = NPX * sizeof(float);
  static_assert(NPX == 64 * 48, "stages are sized for 64 x 48 maps");
  extern __shared__ __align__(128) uint8_t dec_smem_raw[];
  DecodeSmem& sm = *reinterpret_cast<DecodeSmem*>(dec_smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tta = p.maps_flip != nullptr;
  const int L = tta ? 2 : 1;  // stages per map
  const int K = p.num_kpts;
  const int i0 = blockIdx.x * p.items_per_cta;
  const int n_items = max(0, min(p.count, i0 + p.items_per_cta) - i0);
  pdl_launch_dependents();
  if (threadIdx.x < kFullBars) ptx::mbar_init(&sm.full[threadIdx.x], 1);
  if (threadIdx.x < kStages) ptx::mbar_init(&sm.empty[threadIdx.x], 1);
  if (threadIdx.x == 0) {
    sm.q_count = 0;
    sm.next_item = 0;
    sm.tile_mask = (1u << kTiles) - 1u;
  }
  ptx::fence_barrier_init();
  __syncthreads();
  pdl_wait();  // the logits / scalars come from the previous kernels of the stream

  if (warp == kConsumers) {
    if (lane == 0) {
      const int loads = n_items * L;
      for (int n = 0; n < loads; ++n) {
        const int s = n % kStages;
        if (n >= kStages) ptx::mbar_wait(&sm.empty[s], ((n / kStages) & 1) ^ 1);  // the stage's previous map is done with it
        const int item = i0 + (tta ? n >> 1 : n);
        const float* src;
        if (tta && (n & 1)) {
          const int b = item / K, k = item % K;
          src = p.maps_flip + (size_t)(b * K + p.flip_idx[k]) * NPX;
        } else {
          src = p.maps + (size_t)item * NPX;
        }
        uint64_t* bar = &sm.full[n % kFullBars];
        ptx::mbar_arrive_expect_tx(bar, MAP_BYTES);
        bulk_load(sm.stage[s], src, MAP_BYTES, bar);
      }
    }
  } else {
    WarpList& ws = sm.warp[warp];
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&sm.next_item, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= n_items) break;
      const int n0 = t * L;
      const int s1 = n0 % kStages, s2 = (n0 + 1) % kStages;
      PP_T(t_w0);
      ptx::mbar_wait(&sm.full[n0 % kFullBars], (n0 / kFullBars) & 1);
      if (tta) ptx::mbar_wait(&sm.full[(n0 + 1) % kFullBars], ((n0 + 1) / kFullBars) & 1);
      PP_T(t_w1);
      PP_TADD(1, t_w0, t_w1);
      auto release = [&](int which) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.empty[which ? s2 : s1]);
      };
      if (p.probe == 1) {
        release(0);
        if (tta) release(1);
        continue;
      }
      if (!decode_sparse<H, W, LOGITS>(p, i0 + t, ws, lane, sm.stage[s1], sm.stage[s2], &sm.tile[0][0], &sm.tile_mask, release)) {
        if (lane == 0) sm.queue[atomicAdd(&sm.q_count, 1)] = (unsigned short)t;
      }
    }
  }
  __syncthreads();  // every map has been claimed and finished: all bulk copies have landed, the ring is free
  const int nq = sm.q_count;
  if (nq == 0 || warp >= kDenseWarps) return;
  for (int q = 0; q < nq; ++q)
    decode_dense<H, W>(p, i0 + sm.queue[q], sm.stage[0], sm.stage[1], sm.stage[2], sm.red_f, sm.red_i);
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

#ifdef PP_DECODE_TIMING
extern "C" __attribute__((visibility("default"))) int pp_debug_decode_timing(unsigned long long* out8, int reset) {
  using namespace pp;
  PP_CHECK_CUDA(cudaDeviceSynchronize());
  PP_CHECK_CUDA(cudaMemcpyFromSymbol(out8, g_dec_timing, sizeof(g_dec_timing)));
  if (reset) {
    unsigned long long z[8] = {};
    PP_CHECK_CUDA(cudaMemcpyToSymbol(g_dec_timing, z, sizeof(z)));
  }
  return PP_OK;
}
#endif

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
  PP_REQUIRE((reinterpret_cast<uintptr_t>(maps) & 15) == 0 && (reinterpret_cast<uintptr_t>(maps_flip) & 15) == 0, PP_ERR_INVALID,
             "pp_decode: maps / maps_flip must be 16-byte aligned (bulk copies into shared memory)");
  static PerDeviceOnce attr_set;
  auto kern = cfg->input_is_logits ? decode_kernel<64, 48, true> : decode_kernel<64, 48, false>;
  if (attr_set.first()) {  // one persistent CTA per SM with the whole 227 KB of shared memory
    PP_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel<64, 48, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeSmem)));
    PP_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel<64, 48, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeSmem)));
  }
  // contiguous ranges of maps, one CTA per SM (more CTAs only when a range would exceed the dense-path queue)
  const int sms = device_sm_count();
  int64_t ctas = count < sms ? count : sms;
  if ((count + ctas - 1) / ctas > kQueueCap) ctas = (count + kQueueCap - 1) / kQueueCap;
  p.items_per_cta = (int)((count + ctas - 1) / ctas);
  p.probe = getenv("PP_DECODE_PROBE") ? atoi(getenv("PP_DECODE_PROBE")) : 0;
  const int grid = (int)((count + p.items_per_cta - 1) / p.items_per_cta);
  PP_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kDecThreads), sizeof(DecodeSmem), (cudaStream_t)stream, p));
  count_launch();
  PP_CHECK_CUDA(cudaGetLastError());
  return PP_OK;
}
