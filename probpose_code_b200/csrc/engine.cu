// pp_engine: the whole ProbPose forward on one GPU, as a fixed sequence of kernel launches on
// the caller's stream inside one caller-owned workspace (no allocation, no host sync).
//
//   crops (uint8 BGR or normalised fp32) --patchify (+ mirrored pass)--> patch GEMM + pos_embed
//   -> depth x [LN -> qkv GEMM -> attention -> proj GEMM (+x) -> LN -> fc1 GEMM + GELU -> fc2 GEMM (+x)]
//   -> LN -> features (B, gh, gw, D) as a GEMM operand
//   -> heatmap branch: 2 x [4 sub-pixel phase GEMMs of ConvTranspose2d(k4,s2,p1) + BN + ReLU] -> 1x1 conv -> logits
//   -> 4 scalar branches: merged 3x3 conv GEMM + BN -> MaxPool -> ReLU -> ... -> 1x1 conv -> sigmoid / ReLU
//   -> fused decode (decode.cu): sparsemax, flip-TTA merge, OKS convolution, argmax, sub-pixel, record.
//
// Reference call chain replaced: TopdownPoseEstimator.predict (topdown.py:86-126) -> extract_feat
// (base.py:196-210) -> ProbMapHead.predict (probmap_head.py:715-804) -> BaseHead.decode
// (base_head.py:33-86) -> ProbMap.decode (probmap.py:170-220).
#include <string>
#include <unordered_map>
#include <vector>

#include "engine_ops.cuh"

namespace pp {

namespace {

struct Param {
  std::string name;
  int64_t numel;
  size_t off;  // bytes from the workspace base
  int group;   // 0 backbone, 1 head
  bool loaded;
};

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off += (bytes + 1023) & ~size_t(1023);
    return o;
  }
};

const char* kBranches[4] = {"probability", "visibility", "oks", "error"};

}  // namespace

}  // namespace pp

struct pp_engine {
  pp_engine_cfg cfg;
  int prec;
  int gh, gw, tokens, D, FF, heads, dh, depth, K, DC, PK;  // PK = 3 * patch^2
  int max_b2;                                              // 2 * max_batch (flip doubling)
  uint8_t* base;
  size_t bytes;
  std::vector<pp::Param> params;
  std::unordered_map<std::string, int> index;
  bool backbone_ready = false, head_ready = false;
  int64_t last_launches = 0;
  // optional per-class event timing (pp_engine_profile_*)
  bool profiling = false;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  double prof_gemm_flops = 0;
  // the four scalar branches are independent after the shared first convolution: their small
  // pooled-stage convolutions run concurrently on side streams (fork / join by events)
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
  // CUDA-graph replay of the pointer-independent middle of pp_engine_infer (everything between the
  // patch extraction, which reads the caller's crops, and the decode, which writes the caller's records):
  // one captured graph per (batch, passes); see run_body_graphed
  struct Graph { int batch, passes, seen; int64_t launches; cudaGraphExec_t exec; };
  std::vector<Graph> graphs;
  cudaStream_t cap_stream = nullptr;  // capture happens here (the caller's stream may be the legacy default stream)
  int graph_max_images = 0;           // replay when passes * batch <= this; 0 = off, < 0 = no limit
  int64_t graph_replays = 0;
  bool group_phases = true;  // the four phases of a deconvolution as one grouped GEMM launch (PP_NO_GROUPED_GEMM=1: four launches)
  size_t l2_setaside = 0;  // persisting-L2 carve-out available for the residual stream (0: off)
  size_t l2_max_window = 0;  // largest access-policy window of the device (informative)
  size_t g_stride = 0;  // bytes between the per-branch tap-gather buffers
  bool branches = false;  // ProbMapHead: the four scalar branches exist (HeatmapHead: heatmap stack only)

  // packed weights (byte offsets)
  size_t w_patch;
  struct Layer { size_t wqkv, wproj, wfc1, wfc2; };
  std::vector<Layer> layers;
  size_t w_dc[2][4], dc_scale[2], dc_shift[2], w_final;
  size_t w_c1, w_c2[4], w_c3[4], c_scale[3], c_shift[3], tail_w, tail_b;
  // activations (byte offsets)
  size_t x, a_op, qkv, h_op, feat_op, feat_f32;
  size_t feat_bytes = 0, d1_bytes = 0;
  size_t g_op, d1_op, d2_op, logits, c_f32, pool_op, scal, pack_tmp;

  template <typename T = void>
  T* at(size_t off) const { return reinterpret_cast<T*>(base + off); }
  const float* P(const std::string& name) const { return at<float>(params[index.at(name)].off); }
};

namespace pp {

static void add_param(pp_engine* e, Bump& bump, const std::string& name, int64_t numel, int group) {
  Param p;
  p.name = name; p.numel = numel; p.group = group; p.loaded = false;
  p.off = bump.take((size_t)numel * sizeof(float));
  e->index[name] = (int)e->params.size();
  e->params.push_back(p);
}

// Lays out parameters, packed weights and activations.  Used both to size the workspace
// (pp_engine_workspace_bytes) and to fill the engine's offsets (pp_engine_create).
static size_t plan(pp_engine* e) {
  const pp_engine_cfg& c = e->cfg;
  e->prec = c.precision;
  e->gh = (c.img_h + 2 * c.patch_pad - c.patch) / c.patch + 1;
  e->gw = (c.img_w + 2 * c.patch_pad - c.patch) / c.patch + 1;
  e->tokens = e->gh * e->gw;
  e->D = c.embed_dim; e->FF = c.ffn_dim; e->heads = c.heads; e->dh = c.heads ? c.embed_dim / c.heads : 0;
  e->depth = c.depth; e->K = c.num_keypoints; e->DC = c.deconv_channels;
  e->branches = c.deconv_channels > 0 && c.head_kind == PP_HEAD_PROBMAP;
  e->PK = 3 * c.patch * c.patch;
  e->max_b2 = 2 * c.max_batch;
  e->params.clear();
  e->index.clear();
  const int D = e->D, FF = e->FF, DC = e->DC, K = e->K, prec = e->prec;
  const int64_t M = (int64_t)e->tokens * e->max_b2;
  Bump b;

  // ---- raw fp32 parameters, by MMPose state_dict name ----
  if (e->depth > 0) {
    add_param(e, b, "backbone.patch_embed.projection.weight", (int64_t)D * e->PK, 0);
    add_param(e, b, "backbone.patch_embed.projection.bias", D, 0);
    add_param(e, b, "backbone.pos_embed", (int64_t)e->tokens * D, 0);
    for (int l = 0; l < e->depth; ++l) {
      const std::string p = "backbone.layers." + std::to_string(l) + ".";
      add_param(e, b, p + "ln1.weight", D, 0); add_param(e, b, p + "ln1.bias", D, 0);
      add_param(e, b, p + "attn.qkv.weight", (int64_t)3 * D * D, 0); add_param(e, b, p + "attn.qkv.bias", 3 * D, 0);
      add_param(e, b, p + "attn.proj.weight", (int64_t)D * D, 0); add_param(e, b, p + "attn.proj.bias", D, 0);
      add_param(e, b, p + "ln2.weight", D, 0); add_param(e, b, p + "ln2.bias", D, 0);
      add_param(e, b, p + "ffn.layers.0.0.weight", (int64_t)FF * D, 0); add_param(e, b, p + "ffn.layers.0.0.bias", FF, 0);
      add_param(e, b, p + "ffn.layers.1.weight", (int64_t)D * FF, 0); add_param(e, b, p + "ffn.layers.1.bias", D, 0);
    }
    add_param(e, b, "backbone.ln1.weight", D, 0); add_param(e, b, "backbone.ln1.bias", D, 0);
  }
  if (DC > 0) {
    int cin = D;
    for (int i = 0; i < 2; ++i) {
      add_param(e, b, "head.deconv_layers." + std::to_string(3 * i) + ".weight", (int64_t)cin * DC * 16, 1);
      const std::string bn = "head.deconv_layers." + std::to_string(3 * i + 1) + ".";
      for (const char* s : {"weight", "bias", "running_mean", "running_var"}) add_param(e, b, bn + s, DC, 1);
      cin = DC;
    }
    add_param(e, b, "head.final_layer.weight", (int64_t)K * DC, 1);
    add_param(e, b, "head.final_layer.bias", K, 1);
    for (int br = 0; br < (e->branches ? 4 : 0); ++br) {
      const std::string p = std::string("head.") + kBranches[br] + "_layers.";
      for (int j = 0; j < 3; ++j) {
        add_param(e, b, p + std::to_string(4 * j) + ".weight", (int64_t)D * D * 9, 1);
        add_param(e, b, p + std::to_string(4 * j) + ".bias", D, 1);
        for (const char* s : {"weight", "bias", "running_mean", "running_var"})
          add_param(e, b, p + std::to_string(4 * j + 1) + "." + s, D, 1);
      }
      add_param(e, b, p + "12.weight", (int64_t)K * D, 1);
      add_param(e, b, p + "12.bias", K, 1);
    }
  }

  // ---- packed weights ----
  if (e->depth > 0) {
    e->w_patch = b.take(pp_operand_bytes(prec, D, e->PK));
    e->layers.resize(e->depth);
    for (auto& L : e->layers) {
      L.wqkv = b.take(pp_operand_bytes(prec, 3 * D, D));
      L.wproj = b.take(pp_operand_bytes(prec, D, D));
      L.wfc1 = b.take(pp_operand_bytes(prec, FF, D));
      L.wfc2 = b.take(pp_operand_bytes(prec, D, FF));
    }
  }
  if (DC > 0) {
    for (int i = 0; i < 2; ++i) {
      const int cin = i == 0 ? D : DC;
      for (int ph = 0; ph < 4; ++ph) e->w_dc[i][ph] = b.take(pp_operand_bytes(prec, DC, 4 * cin));
      e->dc_scale[i] = b.take(DC * sizeof(float));
      e->dc_shift[i] = b.take(DC * sizeof(float));
    }
    e->w_final = b.take(pp_operand_bytes(prec, K, DC));
    size_t tmp_floats = (size_t)DC * 4 * (D > DC ? D : DC);  // one deconv phase matrix
    if (e->branches) {
      e->w_c1 = b.take(pp_operand_bytes(prec, 4 * D, 9 * D));
      for (int br = 0; br < 4; ++br) {
        e->w_c2[br] = b.take(pp_operand_bytes(prec, D, 9 * D));
        e->w_c3[br] = b.take(pp_operand_bytes(prec, D, 9 * D));
      }
      for (int j = 0; j < 3; ++j) {
        e->c_scale[j] = b.take((size_t)4 * D * sizeof(float));
        e->c_shift[j] = b.take((size_t)4 * D * sizeof(float));
      }
      e->tail_w = b.take((size_t)4 * K * D * sizeof(float));
      e->tail_b = b.take((size_t)4 * K * sizeof(float));
      if ((size_t)4 * D * 9 * D > tmp_floats) tmp_floats = (size_t)4 * D * 9 * D;
    }
    e->pack_tmp = b.take(tmp_floats * sizeof(float));
  }

  // ---- activations ----
  if (e->depth > 0) {
    e->x = b.take((size_t)M * D * sizeof(float));
    e->a_op = b.take(pp_operand_bytes(prec, M, e->PK > D ? e->PK : D));
    e->qkv = b.take((size_t)M * 3 * D * sizeof(float));
    e->h_op = b.take(pp_operand_bytes(prec, M, FF));
  }
  // feat_op and d1_op are shared-border maps (gh + 1) x (gw + 1) / (2 gh + 1) x (2 gw + 1) (epilogue.cuh pad_geom mode 2): the tap
  // operands of the head's implicit-GEMM deconvolutions / 3x3 convolution (borders zeroed in finalize)
  const int64_t Mp0 = (int64_t)e->max_b2 * (e->gh + 1) * (e->gw + 1), Mp1 = (int64_t)e->max_b2 * (2 * e->gh + 1) * (2 * e->gw + 1);
  e->feat_op = b.take(pp_operand_bytes(prec, Mp0, D));
  e->feat_bytes = pp_operand_bytes(prec, Mp0, D);
  e->feat_f32 = b.take((size_t)M * D * sizeof(float));
  if (DC > 0) {
    e->d1_op = b.take(pp_operand_bytes(prec, Mp1, DC));
    e->d1_bytes = pp_operand_bytes(prec, Mp1, DC);
    e->d2_op = b.take(pp_operand_bytes(prec, 16 * M, DC));
    e->logits = b.take((size_t)e->max_b2 * K * 16 * e->tokens * sizeof(float));
    e->scal = b.take((size_t)e->max_b2 * 4 * K * sizeof(float));
    if (e->branches) {
      e->g_stride = (pp_operand_bytes(prec, (int64_t)16 * e->max_b2, 9 * D) + 1023) & ~size_t(1023);
      e->g_op = b.take(4 * e->g_stride);  // tap gathers of the pooled 4x4 / 2x2 stages, one buffer per branch
      e->c_f32 = b.take((size_t)M * 4 * D * sizeof(float));
      e->pool_op = b.take(pp_operand_bytes(prec, (int64_t)16 * e->max_b2, 4 * D));
    }
  }
  return b.off;
}

static int validate_cfg(const pp_engine_cfg* c) {
  PP_REQUIRE(c != nullptr, PP_ERR_INVALID, "engine cfg must be non-NULL");
  PP_REQUIRE(c->precision >= PP_PREC_FP16X3 && c->precision <= PP_PREC_FP32_SIMT, PP_ERR_INVALID, "bad precision %d",
             c->precision);
  PP_REQUIRE(c->max_batch >= 1 && c->max_batch <= 8192, PP_ERR_INVALID, "max_batch %d outside [1, 8192]", c->max_batch);
  PP_REQUIRE(c->patch > 0 && c->patch % 4 == 0 && c->patch_pad >= 0 && c->img_h >= c->patch && c->img_w >= c->patch,
             PP_ERR_INVALID, "bad image / patch geometry %dx%d patch %d pad %d", c->img_h, c->img_w, c->patch, c->patch_pad);
  PP_REQUIRE(c->depth >= 0 && c->depth <= 64, PP_ERR_INVALID, "depth %d", c->depth);
  PP_REQUIRE(c->embed_dim == 384 || c->embed_dim == 768, PP_ERR_UNSUPPORTED, "embed_dim %d not built (384, 768)",
             c->embed_dim);
  if (c->depth > 0) {
    PP_REQUIRE(c->heads > 0 && c->embed_dim % c->heads == 0 && (c->embed_dim / c->heads == 32 || c->embed_dim / c->heads == 64),
               PP_ERR_UNSUPPORTED, "heads %d: head width must be 32 or 64", c->heads);
    PP_REQUIRE(c->ffn_dim > 0 && c->ffn_dim % 64 == 0, PP_ERR_UNSUPPORTED, "ffn_dim %d must be a multiple of 64", c->ffn_dim);
    PP_REQUIRE((3 * c->patch * c->patch) % 64 == 0, PP_ERR_UNSUPPORTED, "patch %d: 3*patch^2 must be a multiple of 64", c->patch);
    const int tok = ((c->img_h + 2 * c->patch_pad - c->patch) / c->patch + 1) * ((c->img_w + 2 * c->patch_pad - c->patch) / c->patch + 1);
    PP_REQUIRE(c->precision == PP_PREC_FP32_SIMT || attention_mma_supported(tok, c->embed_dim / c->heads), PP_ERR_UNSUPPORTED,
               "tensor-core attention is built for 192 tokens (256x192 crops, patch 16, pad 2); this geometry has %d", tok);
  }
  PP_REQUIRE(c->depth > 0 || c->deconv_channels > 0, PP_ERR_INVALID, "engine has neither a backbone nor a head");
  PP_REQUIRE(c->head_kind == PP_HEAD_PROBMAP || c->head_kind == PP_HEAD_HEATMAP, PP_ERR_INVALID, "bad head_kind %d", c->head_kind);
  PP_REQUIRE(c->head_kind != PP_HEAD_HEATMAP || c->deconv_channels == 0 ||
                 (c->blur_kernel_size >= 3 && c->blur_kernel_size <= 31 && (c->blur_kernel_size & 1)),
             PP_ERR_INVALID, "HeatmapHead engine: blur_kernel_size %d must be odd and in [3, 31]", c->blur_kernel_size);
  if (c->deconv_channels > 0) {
    PP_REQUIRE(c->deconv_channels % 64 == 0, PP_ERR_UNSUPPORTED, "deconv_channels %d must be a multiple of 64",
               c->deconv_channels);
    PP_REQUIRE(c->num_keypoints >= 1 && c->num_keypoints <= PP_MAX_KEYPOINTS, PP_ERR_INVALID, "num_keypoints %d outside [1, %d]",
               c->num_keypoints, PP_MAX_KEYPOINTS);
    const int gh = (c->img_h + 2 * c->patch_pad - c->patch) / c->patch + 1, gw = (c->img_w + 2 * c->patch_pad - c->patch) / c->patch + 1;
    PP_REQUIRE(gh == 16 && gw == 12, PP_ERR_UNSUPPORTED,
               "ProbMapHead pools (4,3),(2,2),(2,2) (probmap_head.py:264) need a 16x12 feature map, got %dx%d", gh, gw);
    PP_REQUIRE(c->temperature > 0.f, PP_ERR_INVALID, "temperature must be > 0");
  }
  return PP_OK;
}

// Calls of at most this many images (passes x batch) replay a captured graph by default (measured: 5-7 % less device
// time and 40x less host time per call up to 16 images, neutral beyond: profiles/r01e_graph_bench.jsonl).
constexpr int kGraphDefaultMaxImages = 16;

static pp_gemm_args gemm_args(const pp_engine* e, int64_t m, int n, int k, const void* a, const void* w) {
  pp_gemm_args g = {};
  g.precision = e->prec; g.m = (int)m; g.n = n; g.k = k; g.a = a; g.w = w;
  g.act = PP_ACT_NONE; g.out_kind = PP_OUT_F32; g.ldd = n;
  return g;
}

// Runs one launcher; when profiling, brackets it with an event pair on the stream.
template <typename F>
static int timed(pp_engine* e, int cls, cudaStream_t st, F&& launch) {
  if (!e->profiling) return launch();
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; ++i) {
    if (e->event_pool.empty()) {
      PP_CHECK_CUDA(cudaEventCreate(&ev[i]));
    } else {
      ev[i] = e->event_pool.back();
      e->event_pool.pop_back();
    }
  }
  PP_CHECK_CUDA(cudaEventRecord(ev[0], st));
  const int rc = launch();
  PP_CHECK_CUDA(cudaEventRecord(ev[1], st));
  e->spans.push_back({cls, ev[0], ev[1]});
  return rc;
}

static int gemm(pp_engine* e, const pp_gemm_args& g, cudaStream_t st) {
  if (e->profiling) e->prof_gemm_flops += 2.0 * g.m * g.n * g.k;
  return timed(e, PP_KC_GEMM, st, [&] { return gemm_dispatch(g, st); });
}

// `count` GEMMs that share everything but W, the tap shifts and the phase: one launch (gemm_dispatch_group).
static int gemm_group(pp_engine* e, const pp_gemm_args* g, int count, cudaStream_t st) {
  if (e->profiling)
    for (int i = 0; i < count; ++i) e->prof_gemm_flops += 2.0 * g[i].m * g[i].n * g[i].k;
  return timed(e, PP_KC_GEMM, st, [&] { return gemm_dispatch_group(g, count, st); });
}

#define PP_TRY(expr)            \
  do {                          \
    const int _rc = (expr);     \
    if (_rc != PP_OK) return _rc; \
  } while (0)

static int to_operand(const pp_engine* e, const float* src, int64_t rows, int64_t k, size_t dst_off, cudaStream_t st) {
  return pp_operand_from_f32(e->prec, src, rows, k, k, e->at<>(dst_off), st);
}

// ---- forward pieces ---------------------------------------------------------------------------
// crops -> patch matrix (the only kernel of the backbone that reads caller memory)
static int run_patchify(pp_engine* e, const uint8_t* u8, const float* xf, int batch, int passes, cudaStream_t st) {
  const int prec = e->prec;
  PatchifyParams pp_;
  pp_.u8_bgr = u8; pp_.x_f32 = xf; pp_.batch = batch; pp_.passes = passes;
  pp_.img_h = e->cfg.img_h; pp_.img_w = e->cfg.img_w; pp_.patch = e->cfg.patch; pp_.pad = e->cfg.patch_pad;
  pp_.gh = e->gh; pp_.gw = e->gw;
  for (int c = 0; c < 3; ++c) { pp_.mean[c] = e->cfg.mean[c]; pp_.inv_std[c] = 1.0f / e->cfg.std[c]; }
  return timed(e, PP_KC_OTHER, st, [&] { return launch_patchify(prec, pp_, e->at<>(e->a_op), st); });
}

static int run_encoder_layers(pp_engine* e, int batch, int passes, bool want_f32, cudaStream_t st);

// patch matrix -> features.  The residual stream stays resident in L2 for the whole encoder (common.cuh L2Window) when
// the engine's largest one fits the carve-out (pp_engine_create).
static int run_encoder(pp_engine* e, int batch, int passes, bool want_f32, cudaStream_t st) {
  const size_t x_bytes = (size_t)passes * batch * e->tokens * e->D * sizeof(float);
  if (e->l2_setaside > 0 && x_bytes > 0) L2Window::set(e->at<>(e->x), x_bytes, 1.0f);  // x_bytes <= the carve-out by construction
  const int rc = run_encoder_layers(e, batch, passes, want_f32, st);
  L2Window::clear();
  return rc;
}

static int run_encoder_layers(pp_engine* e, int batch, int passes, bool want_f32, cudaStream_t st) {
  const int D = e->D, FF = e->FF, prec = e->prec;
  const int64_t M = (int64_t)passes * batch * e->tokens;
  {
    pp_gemm_args g = gemm_args(e, M, D, e->PK, e->at<>(e->a_op), e->at<>(e->w_patch));
    g.shift = e->P("backbone.patch_embed.projection.bias");
    g.residual = e->P("backbone.pos_embed"); g.res_mod = e->tokens;
    g.d = e->at<>(e->x);
    PP_TRY(gemm(e, g, st));
  }
  float* x = e->at<float>(e->x);
  for (int l = 0; l < e->depth; ++l) {
    const std::string p = "backbone.layers." + std::to_string(l) + ".";
    const pp_engine::Layer& L = e->layers[l];
    PP_TRY(timed(e, PP_KC_OTHER, st, [&] { return launch_layernorm(prec, x, e->P(p + "ln1.weight"), e->P(p + "ln1.bias"), e->cfg.ln_eps, M, D, e->at<>(e->a_op),
                            nullptr, st); }));
    pp_gemm_args g = gemm_args(e, M, 3 * D, D, e->at<>(e->a_op), e->at<>(L.wqkv));
    g.shift = e->P(p + "attn.qkv.bias"); g.d = e->at<>(e->qkv);
    const bool tc_attn = prec != PP_PREC_FP32_SIMT;  // q, k, v leave the GEMM pre-split for the tensor-core attention
    if (tc_attn) { g.out_kind = PP_OUT_OPERAND; g.ldd = 3 * D; }
    PP_TRY(gemm(e, g, st));
    PP_TRY(timed(e, PP_KC_ATTENTION, st, [&] {
      return tc_attn ? (attention_use_tc() ? launch_attention_tc : launch_attention_mma)(prec, e->at<>(e->qkv), passes * batch, e->tokens, e->heads, e->dh, e->at<>(e->a_op), st)
                     : launch_attention(prec, e->at<float>(e->qkv), passes * batch, e->tokens, e->heads, e->dh, e->at<>(e->a_op), st);
    }));
    g = gemm_args(e, M, D, D, e->at<>(e->a_op), e->at<>(L.wproj));
    g.shift = e->P(p + "attn.proj.bias"); g.residual = x; g.d = x;
    PP_TRY(gemm(e, g, st));
    PP_TRY(timed(e, PP_KC_OTHER, st, [&] { return launch_layernorm(prec, x, e->P(p + "ln2.weight"), e->P(p + "ln2.bias"), e->cfg.ln_eps, M, D, e->at<>(e->a_op),
                            nullptr, st); }));
    g = gemm_args(e, M, FF, D, e->at<>(e->a_op), e->at<>(L.wfc1));
    g.shift = e->P(p + "ffn.layers.0.0.bias"); g.act = PP_ACT_GELU; g.out_kind = PP_OUT_OPERAND; g.ldd = FF;
    g.d = e->at<>(e->h_op);
    PP_TRY(gemm(e, g, st));
    g = gemm_args(e, M, D, FF, e->at<>(e->h_op), e->at<>(L.wfc2));
    g.shift = e->P(p + "ffn.layers.1.bias"); g.residual = x; g.d = x;
    PP_TRY(gemm(e, g, st));
  }
  PP_TRY(timed(e, PP_KC_OTHER, st, [&] { return launch_layernorm(prec, x, e->P("backbone.ln1.weight"), e->P("backbone.ln1.bias"), e->cfg.ln_eps, M, D,
                          e->at<>(e->feat_op), want_f32 ? e->at<float>(e->feat_f32) : nullptr, st, e->gh, e->gw); }));
  return PP_OK;
}

// feat_op (n_img * tokens, D) -> logits (n_img, K, 16 * tokens), scalars (n_img, 4, K)
static int run_head(pp_engine* e, int n_img, float* logits, float* scalars, cudaStream_t st) {
  const int D = e->D, DC = e->DC, K = e->K, prec = e->prec;
  const int64_t M = (int64_t)n_img * e->tokens;
  // --- heatmap branch: two stride-2 deconvs as 4 sub-pixel phase GEMMs each, then the 1x1 conv.
  // Implicit GEMM: the A operand is the zero-bordered input map itself, read at 4 row shifts (taps).
  for (int i = 0; i < 2; ++i) {
    const int cin = i == 0 ? D : DC, h = e->gh << i, w = e->gw << i;
    const int64_t rows = (int64_t)n_img * (h + 1) * (w + 1);
    const void* src = i == 0 ? e->at<>(e->feat_op) : e->at<>(e->d1_op);
    void* dst = i == 0 ? e->at<>(e->d1_op) : e->at<>(e->d2_op);
    pp_gemm_args phases[4];  // the four sub-pixel phases: one grouped launch
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      pp_gemm_args& g = phases[ph];
      g = gemm_args(e, rows, DC, 4 * cin, src, e->at<>(e->w_dc[i][ph]));
      g.a_taps = 4;
      for (int t = 0; t < 4; ++t) {
        int dy, dx, kk;
        deconv_tap(py, t >> 1, &dy, &kk);
        deconv_tap(px, t & 1, &dx, &kk);
        g.a_tap_shift[t] = dy * (w + 1) + dx;
      }
      g.scale = e->at<float>(e->dc_scale[i]); g.shift = e->at<float>(e->dc_shift[i]); g.act = PP_ACT_RELU;
      g.out_kind = PP_OUT_OPERAND; g.ldd = DC; g.d = dst;
      g.in_pad = 2; g.in_h = h; g.in_w = w; g.out_pad = i == 0 ? 2 : 0;
      g.up_hin = h; g.up_win = w; g.up_py = py; g.up_px = px;
      if (e->profiling) e->prof_gemm_flops -= 2.0 * (rows - (double)n_img * h * w) * DC * 4 * cin;  // border rows are not algorithmic work
    }
    if (e->group_phases) {
      PP_TRY(gemm_group(e, phases, 4, st));
    } else {
      for (int ph = 0; ph < 4; ++ph) PP_TRY(gemm(e, phases[ph], st));
    }
  }
  {
    pp_gemm_args g = gemm_args(e, 16 * M, K, DC, e->at<>(e->d2_op), e->at<>(e->w_final));
    g.shift = e->P("head.final_layer.bias"); g.out_kind = PP_OUT_PLANES; g.plane = 16 * e->tokens; g.d = logits;
    PP_TRY(gemm(e, g, st));
  }
  if (!e->branches) return PP_OK;  // HeatmapHead: the heatmap stack is the whole head
  // --- four scalar branches; the first conv of all four shares its input -> one GEMM, N = 4 D,
  // 9 taps over the zero-bordered feature map ---
  {
    const int64_t rows = (int64_t)n_img * (e->gh + 1) * (e->gw + 1);
    pp_gemm_args g = gemm_args(e, rows, 4 * D, 9 * D, e->at<>(e->feat_op), e->at<>(e->w_c1));
    g.a_taps = 9;
    for (int t = 0; t < 9; ++t) g.a_tap_shift[t] = (t / 3 - 1) * (e->gw + 1) + (t % 3 - 1);
    g.in_pad = 2; g.in_h = e->gh; g.in_w = e->gw;
    g.scale = e->at<float>(e->c_scale[0]); g.shift = e->at<float>(e->c_shift[0]); g.d = e->at<>(e->c_f32);
    if (e->profiling) e->prof_gemm_flops -= 2.0 * (rows - (double)M) * 4 * D * 9 * D;
    PP_TRY(gemm(e, g, st));
  }
  GatherParams g3 = {};
  g3.ntaps = 9;
  for (int t = 0; t < 9; ++t) { g3.dy[t] = t / 3 - 1; g3.dx[t] = t % 3 - 1; }
  g3.batch = n_img; g3.c = D;
  int h = e->gh, w = e->gw;
  const int pool[3][2] = {{4, 3}, {2, 2}, {2, 2}};  // probmap_head.py:264
  for (int j = 1; j < 3; ++j) {
    PP_TRY(timed(e, PP_KC_OTHER, st, [&] { return launch_pool_relu(prec, e->at<float>(e->c_f32), n_img, h, w, 4 * D, pool[j - 1][0], pool[j - 1][1],
                            e->at<>(e->pool_op), st); }));
    h /= pool[j - 1][0]; w /= pool[j - 1][1];
    const int64_t rows = (int64_t)n_img * h * w;
    if (e->fork_ev) PP_CHECK_CUDA(cudaEventRecord(e->fork_ev, st));
    for (int br = 0; br < 4; ++br) {
      cudaStream_t bs = (br == 0 || !e->fork_ev) ? st : e->side[br - 1];
      if (bs != st) PP_CHECK_CUDA(cudaStreamWaitEvent(bs, e->fork_ev, 0));
      void* gbuf = e->at<>(e->g_op + br * e->g_stride);
      g3.h = h; g3.w = w; g3.src_c = 4 * D; g3.c_off = br * D;
      PP_TRY(timed(e, PP_KC_OTHER, bs, [&] { return launch_gather_taps(prec, g3, e->at<>(e->pool_op), gbuf, bs); }));
      pp_gemm_args g = gemm_args(e, rows, D, 9 * D, gbuf, e->at<>(j == 1 ? e->w_c2[br] : e->w_c3[br]));
      g.scale = e->at<float>(e->c_scale[j]) + br * D; g.shift = e->at<float>(e->c_shift[j]) + br * D;
      g.ldd = 4 * D; g.d = e->at<float>(e->c_f32) + br * D;
      // four of these run at once (side streams): pp_gemm's width choice for a GEMM alone (64 / 32 columns, to spread a long
      // K loop over idle SMs) over-subscribes the SMs 2.6-fold; 128 / 64 keep the MMAs wide (A/B: 10 389 / 10 389 persons/s
      // against 10 367 / 10 252 with the automatic widths at 64 crops with the flipped pass)
      if (e->fork_ev) {  // the widest tile that still gives the four GEMMs together half an SM count of tiles
        const int64_t mt = (rows + 127) / 128;
        g.tile_n = 32;
        for (int bn = 128; bn > 32; bn >>= 1)
          if (4 * mt * ((D + bn - 1) / bn) >= device_sm_count() / 2) { g.tile_n = bn; break; }
      }
      PP_TRY(gemm(e, g, bs));
      if (bs != st) {
        PP_CHECK_CUDA(cudaEventRecord(e->join_ev[br - 1], bs));
        PP_CHECK_CUDA(cudaStreamWaitEvent(st, e->join_ev[br - 1], 0));
      }
    }
  }
  PP_TRY(timed(e, PP_KC_OTHER, st, [&] { return launch_branch_tail(e->at<float>(e->c_f32), n_img, D, K, e->at<float>(e->tail_w), e->at<float>(e->tail_b), scalars, st); }));
  return PP_OK;
}

static int run_backbone(pp_engine* e, const uint8_t* u8, const float* xf, int batch, int passes, bool want_f32,
                        cudaStream_t st) {
  PP_TRY(run_patchify(e, u8, xf, batch, passes, st));
  return run_encoder(e, batch, passes, want_f32, st);
}

// The middle of pp_engine_infer: patch matrix -> logits + branch scalars, all inside the workspace.
static int run_body(pp_engine* e, int batch, int passes, cudaStream_t st) {
  PP_TRY(run_encoder(e, batch, passes, false, st));
  return run_head(e, passes * batch, e->at<float>(e->logits), e->at<float>(e->scal), st);
}

static void graph_give_up(pp_engine* e) {  // capture is unavailable here: plain launches from now on
  e->graph_max_images = 0;
  cudaGetLastError();
}

// run_body through a CUDA graph: the first call of a (batch, passes) shape runs plainly (and finishes the
// one-time function-attribute set-up of every launcher), the second one is captured on the engine's own
// stream - kernel nodes keep their programmatic-dependent-launch edges, the branch side streams become
// parallel branches of the graph - and every later call is ONE cudaGraphLaunch on the caller's stream.
static int run_body_graphed(pp_engine* e, int batch, int passes, cudaStream_t st) {
  const int64_t images = (int64_t)batch * passes;
  if (e->profiling || e->graph_max_images == 0 || (e->graph_max_images > 0 && images > e->graph_max_images))
    return run_body(e, batch, passes, st);
  pp_engine::Graph* g = nullptr;
  for (auto& it : e->graphs)
    if (it.batch == batch && it.passes == passes) g = &it;
  if (!g) {
    if (e->graphs.size() >= 64) {  // a caller cycling through many shapes: drop the oldest graph
      if (e->graphs.front().exec) cudaGraphExecDestroy(e->graphs.front().exec);
      e->graphs.erase(e->graphs.begin());
    }
    e->graphs.push_back({batch, passes, 1, 0, nullptr});
    return run_body(e, batch, passes, st);
  }
  if (!g->exec) {
    if (!e->cap_stream && cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
      graph_give_up(e);
      return run_body(e, batch, passes, st);
    }
    if (cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      graph_give_up(e);
      return run_body(e, batch, passes, st);
    }
    const int64_t c0 = g_launch_count;
    const int rc = run_body(e, batch, passes, e->cap_stream);
    const int64_t launches = g_launch_count - c0;
    g_launch_count = c0;  // nothing has run yet: the replay below counts them
    cudaGraph_t graph = nullptr;
    cudaError_t err = cudaStreamEndCapture(e->cap_stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == PP_OK && err == cudaSuccess && graph) err = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != PP_OK || err != cudaSuccess || !exec) {
      graph_give_up(e);
      return run_body(e, batch, passes, st);
    }
    g->exec = exec;
    g->launches = launches;
  }
  PP_CHECK_CUDA(cudaGraphLaunch(g->exec, st));
  count_launch((int)g->launches);
  ++e->graph_replays;
  return PP_OK;
}

static int check_batch(const pp_engine* e, int batch, int passes, const char* what) {
  PP_REQUIRE(e != nullptr, PP_ERR_INVALID, "%s: engine is NULL", what);
  PP_REQUIRE(batch >= 0 && (int64_t)batch * passes <= e->max_b2, PP_ERR_INVALID,
             "%s: batch %d (x%d passes) exceeds the engine's max_batch %d", what, batch, passes, e->cfg.max_batch);
  return PP_OK;
}

}  // namespace pp

// =================================== C ABI ====================================================
using namespace pp;

extern "C" size_t pp_engine_workspace_bytes(const pp_engine_cfg* cfg) {
  if (validate_cfg(cfg) != PP_OK) return 0;
  pp_engine tmp;
  tmp.cfg = *cfg;
  return plan(&tmp);
}

extern "C" int pp_engine_create(const pp_engine_cfg* cfg, void* workspace, size_t workspace_bytes, pp_engine** out) {
  PP_REQUIRE(out != nullptr, PP_ERR_INVALID, "pp_engine_create: out is NULL");
  *out = nullptr;
  PP_TRY(validate_cfg(cfg));
  PP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, PP_ERR_INVALID,
             "pp_engine_create: workspace must be non-NULL and 1024-byte aligned");
  pp_engine* e = new pp_engine();
  e->cfg = *cfg;
  const size_t need = plan(e);
  if (workspace_bytes < need) {
    set_error("pp_engine_create: workspace of %zu bytes is smaller than the %zu required", workspace_bytes, need);
    delete e;
    return PP_ERR_INVALID;
  }
  e->base = reinterpret_cast<uint8_t*>(workspace);
  e->bytes = workspace_bytes;
  if (e->branches && getenv("PP_NO_BRANCH_STREAMS") == nullptr) {  // side streams of the scalar branches (no device memory involved)
    bool ok = cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && ok; ++i)
      ok = cudaStreamCreateWithFlags(&e->side[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&e->join_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      set_error("pp_engine_create: could not create the branch streams: %s", cudaGetErrorString(cudaGetLastError()));
      pp_engine_destroy(e);
      return PP_ERR_CUDA;
    }
  }
  if (e->depth > 0 && getenv("PP_NO_L2_PERSIST") == nullptr) {
    // L2 carve-out for the residual stream.  Only for engines whose largest residual stream fits a third of the L2
    // (<= 64 crops with the flipped pass on B200): a partly persisting stream or a carve-out that leaves the other
    // activations too little cache loses more than it gains (256 crops per call: 11 016 -> 8 694 persons/s with a
    // 29 % persisting window).  The limit is device-wide and is only ever raised.
    int dev = 0, max_persist = 0, max_window = 0, l2 = 0;
    size_t cur = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess && max_persist > 0 &&
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev) == cudaSuccess && max_window > 0 &&
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev) == cudaSuccess && l2 > 0) {
      const size_t x_max = (size_t)e->max_b2 * e->tokens * e->D * sizeof(float);
      size_t cap = (size_t)l2 / 3 < (size_t)max_persist ? (size_t)l2 / 3 : (size_t)max_persist;
      const char* mb = getenv("PP_L2_PERSIST_MB");
      if (mb && ((size_t)atoi(mb) << 20) < cap) cap = (size_t)atoi(mb) << 20;
      if (x_max <= cap && x_max <= (size_t)max_window) {
        if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) != cudaSuccess) cur = 0;
        if (x_max > cur && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, x_max) == cudaSuccess) cur = x_max;
        if (cur >= x_max) { e->l2_setaside = x_max; e->l2_max_window = (size_t)max_window; }
      }
    }
    cudaGetLastError();  // an unsupported limit is not an error of the engine
  }
  e->group_phases = getenv("PP_NO_GROUPED_GEMM") == nullptr;
  {  // graph replay of small calls is on by default; PP_ENGINE_GRAPH = 0 (off), -1 (every size) or a limit in images
    const char* env = getenv("PP_ENGINE_GRAPH");
    e->graph_max_images = env ? atoi(env) : kGraphDefaultMaxImages;
  }
  *out = e;
  return PP_OK;
}

extern "C" int pp_engine_set_graph(pp_engine* e, int32_t max_images) {
  PP_REQUIRE(e != nullptr, PP_ERR_INVALID, "pp_engine_set_graph: engine is NULL");
  e->graph_max_images = max_images;
  return PP_OK;
}

extern "C" int64_t pp_engine_graph_replay_count(const pp_engine* e) { return e ? e->graph_replays : 0; }

extern "C" void pp_engine_destroy(pp_engine* e) {
  if (!e) return;
  for (auto& g : e->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  for (auto& s : e->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  for (auto ev : e->event_pool) cudaEventDestroy(ev);
  if (e->fork_ev) cudaEventDestroy(e->fork_ev);
  for (int i = 0; i < 3; ++i) {
    if (e->join_ev[i]) cudaEventDestroy(e->join_ev[i]);
    if (e->side[i]) cudaStreamDestroy(e->side[i]);
  }
  delete e;
}

extern "C" int pp_engine_load(pp_engine* e, const char* name, const float* data, int64_t numel, void* stream) {
  PP_REQUIRE(e && name && data, PP_ERR_INVALID, "pp_engine_load: NULL argument");
  const std::string n(name);
  if (n.size() >= 19 && n.compare(n.size() - 19, 19, "num_batches_tracked") == 0) return PP_OK;  // BN bookkeeping, unused
  auto it = e->index.find(n);
  PP_REQUIRE(it != e->index.end(), PP_ERR_INVALID, "pp_engine_load: unknown parameter '%s'", name);
  Param& p = e->params[it->second];
  PP_REQUIRE(p.numel == numel, PP_ERR_INVALID, "pp_engine_load: '%s' has %lld elements, expected %lld", name,
             (long long)numel, (long long)p.numel);
  PP_CHECK_CUDA(cudaMemcpyAsync(e->base + p.off, data, (size_t)numel * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  p.loaded = true;
  e->backbone_ready = e->head_ready = false;  // finalize again
  return PP_OK;
}

extern "C" int pp_engine_finalize(pp_engine* e, void* stream) {
  PP_REQUIRE(e != nullptr, PP_ERR_INVALID, "pp_engine_finalize: engine is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int loaded[2] = {0, 0}, total[2] = {0, 0};
  const Param* missing[2] = {nullptr, nullptr};
  for (const Param& p : e->params) {
    ++total[p.group];
    if (p.loaded) ++loaded[p.group];
    else if (!missing[p.group]) missing[p.group] = &p;
  }
  for (int g = 0; g < 2; ++g)
    PP_REQUIRE(loaded[g] == 0 || loaded[g] == total[g], PP_ERR_STATE,
               "pp_engine_finalize: %d of %d %s parameters loaded; first missing '%s'", loaded[g], total[g],
               g ? "head" : "backbone", missing[g]->name.c_str());
  PP_REQUIRE(loaded[0] + loaded[1] > 0, PP_ERR_STATE, "pp_engine_finalize: no parameters loaded");
  PP_CHECK_CUDA(cudaMemsetAsync(e->at<>(e->feat_op), 0, e->feat_bytes, st));  // zero border of the padded feature map
  const int D = e->D, FF = e->FF, DC = e->DC, K = e->K;
  if (total[0] > 0 && loaded[0] == total[0]) {
    PP_TRY(to_operand(e, e->P("backbone.patch_embed.projection.weight"), D, e->PK, e->w_patch, st));
    for (int l = 0; l < e->depth; ++l) {
      const std::string p = "backbone.layers." + std::to_string(l) + ".";
      PP_TRY(to_operand(e, e->P(p + "attn.qkv.weight"), 3 * D, D, e->layers[l].wqkv, st));
      PP_TRY(to_operand(e, e->P(p + "attn.proj.weight"), D, D, e->layers[l].wproj, st));
      PP_TRY(to_operand(e, e->P(p + "ffn.layers.0.0.weight"), FF, D, e->layers[l].wfc1, st));
      PP_TRY(to_operand(e, e->P(p + "ffn.layers.1.weight"), D, FF, e->layers[l].wfc2, st));
    }
    e->backbone_ready = true;
  }
  if (total[1] > 0 && loaded[1] == total[1]) {
    float* tmp = e->at<float>(e->pack_tmp);
    for (int i = 0; i < 2; ++i) {
      const int cin = i == 0 ? D : DC;
      const std::string w = "head.deconv_layers." + std::to_string(3 * i) + ".weight";
      const std::string bn = "head.deconv_layers." + std::to_string(3 * i + 1) + ".";
      for (int ph = 0; ph < 4; ++ph) {
        PP_TRY(launch_pack_deconv_phase(e->P(w), cin, DC, ph >> 1, ph & 1, tmp, st));
        PP_TRY(to_operand(e, tmp, DC, 4 * cin, e->w_dc[i][ph], st));
      }
      PP_TRY(launch_fold_bn(e->P(bn + "weight"), e->P(bn + "bias"), e->P(bn + "running_mean"), e->P(bn + "running_var"),
                            nullptr, e->cfg.bn_eps, DC, e->at<float>(e->dc_scale[i]), e->at<float>(e->dc_shift[i]), st));
    }
    PP_TRY(to_operand(e, e->P("head.final_layer.weight"), K, DC, e->w_final, st));
    for (int br = 0; br < (e->branches ? 4 : 0); ++br) {
      const std::string p = std::string("head.") + kBranches[br] + "_layers.";
      PP_TRY(launch_pack_conv3x3(e->P(p + "0.weight"), D, D, tmp + (size_t)br * D * 9 * D, st));
      for (int j = 0; j < 3; ++j) {
        const std::string bn = p + std::to_string(4 * j + 1) + ".";
        PP_TRY(launch_fold_bn(e->P(bn + "weight"), e->P(bn + "bias"), e->P(bn + "running_mean"), e->P(bn + "running_var"),
                              e->P(p + std::to_string(4 * j) + ".bias"), e->cfg.bn_eps, D,
                              e->at<float>(e->c_scale[j]) + br * D, e->at<float>(e->c_shift[j]) + br * D, st));
      }
      PP_CHECK_CUDA(cudaMemcpyAsync(e->at<float>(e->tail_w) + (size_t)br * K * D, e->P(p + "12.weight"),
                                    (size_t)K * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
      PP_CHECK_CUDA(cudaMemcpyAsync(e->at<float>(e->tail_b) + (size_t)br * K, e->P(p + "12.bias"), (size_t)K * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
    }
    if (e->branches) PP_TRY(to_operand(e, tmp, 4 * D, 9 * D, e->w_c1, st));
    for (int br = 0; br < (e->branches ? 4 : 0); ++br) {
      const std::string p = std::string("head.") + kBranches[br] + "_layers.";
      PP_TRY(launch_pack_conv3x3(e->P(p + "4.weight"), D, D, tmp, st));
      PP_TRY(to_operand(e, tmp, D, 9 * D, e->w_c2[br], st));
      PP_TRY(launch_pack_conv3x3(e->P(p + "8.weight"), D, D, tmp, st));
      PP_TRY(to_operand(e, tmp, D, 9 * D, e->w_c3[br], st));
    }
    PP_CHECK_CUDA(cudaMemsetAsync(e->at<>(e->d1_op), 0, e->d1_bytes, st));
    e->head_ready = true;
  }
  return PP_OK;
}

extern "C" int pp_engine_backbone(pp_engine* e, const float* x, int32_t batch, float* feat_nchw, void* stream) {
  PP_TRY(check_batch(e, batch, 1, "pp_engine_backbone"));
  PP_REQUIRE(e->backbone_ready, PP_ERR_STATE, "pp_engine_backbone: backbone weights not loaded / finalized");
  PP_REQUIRE(batch == 0 || (x && feat_nchw), PP_ERR_INVALID, "pp_engine_backbone: NULL tensor");
  const int64_t before = g_launch_count;
  if (batch > 0) {
    PP_TRY(run_backbone(e, nullptr, x, batch, 1, true, (cudaStream_t)stream));
    PP_TRY(launch_rows_to_nchw(e->at<float>(e->feat_f32), batch, e->tokens, e->D, feat_nchw, (cudaStream_t)stream));
  }
  e->last_launches = g_launch_count - before;
  return PP_OK;
}

extern "C" int pp_engine_head(pp_engine* e, const float* feat_nchw, int32_t batch, float* heat_logits, float* scalars,
                              void* stream) {
  PP_TRY(check_batch(e, batch, 1, "pp_engine_head"));
  PP_REQUIRE(e->head_ready, PP_ERR_STATE, "pp_engine_head: head weights not loaded / finalized");
  PP_REQUIRE(batch == 0 || (feat_nchw && heat_logits && (scalars || !e->branches)), PP_ERR_INVALID, "pp_engine_head: NULL tensor");
  const int64_t before = g_launch_count;
  if (batch > 0) {
    PP_TRY(launch_nchw_to_operand(e->prec, feat_nchw, batch, e->tokens, e->D, e->at<>(e->feat_op), (cudaStream_t)stream, e->gh, e->gw));
    PP_TRY(run_head(e, batch, heat_logits, scalars, (cudaStream_t)stream));
  }
  e->last_launches = g_launch_count - before;
  return PP_OK;
}

extern "C" int pp_engine_infer(pp_engine* e, const uint8_t* crops_u8_bgr, const float* x_f32, int32_t batch,
                               int32_t flip_test, const int32_t* flip_indices, float* records, float* merged_out,
                               void* stream) {
  const int passes = flip_test ? 2 : 1;
  PP_TRY(check_batch(e, batch, passes, "pp_engine_infer"));
  PP_REQUIRE(e->backbone_ready && e->head_ready, PP_ERR_STATE,
             "pp_engine_infer: needs backbone and head weights loaded and finalized");
  PP_REQUIRE(batch == 0 || records, PP_ERR_INVALID, "pp_engine_infer: records is NULL");
  PP_REQUIRE(!flip_test || flip_indices, PP_ERR_INVALID, "pp_engine_infer: flip_test needs flip_indices");
  const int64_t before = g_launch_count;
  if (batch > 0) {
    cudaStream_t st = (cudaStream_t)stream;
    PP_TRY(run_patchify(e, crops_u8_bgr, x_f32, batch, passes, st));
    float* logits = e->at<float>(e->logits);
    float* scal = e->at<float>(e->scal);
    PP_TRY(run_body_graphed(e, batch, passes, st));
    if (e->cfg.head_kind == PP_HEAD_HEATMAP) {  // HeatmapHead + UDPHeatmap: records are (B, K, 3)
      pp_udp_cfg uc;
      uc.num_keypoints = e->K; uc.height = 4 * e->gh; uc.width = 4 * e->gw; uc.blur_kernel_size = e->cfg.blur_kernel_size;
      const size_t stride = (size_t)batch * e->K * 16 * e->tokens;
      PP_TRY(timed(e, PP_KC_DECODE, st, [&] {
        return pp_decode_udp(&uc, logits, flip_test ? logits + stride : nullptr, flip_indices, batch, records, merged_out, stream);
      }));
      e->last_launches = g_launch_count - before;
      return PP_OK;
    }
    pp_decode_cfg dc;
    dc.num_keypoints = e->K; dc.height = 4 * e->gh; dc.width = 4 * e->gw; dc.input_is_logits = 1;
    dc.temperature = e->cfg.temperature; dc.normalize = e->cfg.normalize; dc.error_divisor = 0.f;
    const size_t map_stride = (size_t)batch * e->K * 16 * e->tokens, sc_stride = (size_t)batch * 4 * e->K;
    PP_TRY(timed(e, PP_KC_DECODE, st, [&] {
      return pp_decode(&dc, logits, flip_test ? logits + map_stride : nullptr, flip_indices, scal,
                       flip_test ? scal + sc_stride : nullptr, batch, records, merged_out, stream);
    }));
  }
  e->last_launches = g_launch_count - before;
  return PP_OK;
}

extern "C" int64_t pp_engine_last_launch_count(const pp_engine* e) { return e ? e->last_launches : 0; }

extern "C" int pp_engine_profile_begin(pp_engine* e) {
  PP_REQUIRE(e != nullptr, PP_ERR_INVALID, "pp_engine_profile_begin: engine is NULL");
  for (auto& s : e->spans) { e->event_pool.push_back(s.a); e->event_pool.push_back(s.b); }
  e->spans.clear();
  e->prof_gemm_flops = 0;
  e->profiling = true;
  return PP_OK;
}

extern "C" int pp_engine_profile_end(pp_engine* e, pp_profile* out, void* stream) {
  PP_REQUIRE(e && out, PP_ERR_INVALID, "pp_engine_profile_end: NULL argument");
  PP_REQUIRE(e->profiling, PP_ERR_STATE, "pp_engine_profile_end without pp_engine_profile_begin");
  e->profiling = false;
  PP_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  for (int c = 0; c < PP_KC_COUNT; ++c) { out->ms[c] = 0; out->launches[c] = 0; }
  for (auto& s : e->spans) {
    float ms = 0.f;
    PP_CHECK_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
    out->ms[s.cls] += ms;
    out->launches[s.cls] += 1;
    e->event_pool.push_back(s.a);
    e->event_pool.push_back(s.b);
  }
  e->spans.clear();
  out->gemm_flops = e->prof_gemm_flops;
  return PP_OK;
}
