// Launchers of the non-GEMM kernels the engine strings together (vit_ops.cu, head_ops.cu).
// All of them take operands in the engine precision (see common.cuh "GEMM operand formats")
// and return a pp_status.
#pragma once

#include "common.cuh"

namespace pp {

// ---- vit_ops.cu ---------------------------------------------------------------------------
struct PatchifyParams {
  const uint8_t* u8_bgr;  // (B, 3, H, W) uint8 BGR, or NULL
  const float* x_f32;     // (B, 3, H, W) fp32 normalised RGB, or NULL
  int batch;              // source crops
  int passes;             // 1, or 2: rows of pass 1 are the crops mirrored left-right
  int img_h, img_w, patch, pad, gh, gw;
  float mean[3], inv_std[3];  // RGB order
};
// -> operand (passes * B * gh * gw, 3 * patch * patch), k = c * patch^2 + ky * patch + kx
int launch_patchify(int prec, const PatchifyParams& p, void* a_op, cudaStream_t st);

// Row LayerNorm of fp32 (rows, d) -> operand (rows, d) and, optionally, fp32 (rows, d).
// pad_gw > 0: operand rows go to the interior of a shared-border (pad_gh + 1) x (pad_gw + 1) map per image
// (epilogue.cuh pad_geom mode 2).
int launch_layernorm(int prec, const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d,
                     void* out_op, float* out_f32, cudaStream_t st, int pad_gh = 0, int pad_gw = 0);

// Global multi-head attention on packed qkv fp32 (B * n, 3 * heads * dh) -> operand (B * n, heads * dh).
int launch_attention(int prec, const float* qkv, int batch, int n, int heads, int dh, void* out_op, cudaStream_t st);

// Tensor-core attention (attention.cu): qkv_op is the qkv GEMM's PP_OUT_OPERAND output
// (B * n, 3 * heads * dh) in `prec`; out_op the proj GEMM's A operand.
bool attention_mma_supported(int n, int dh);
int launch_attention_mma(int prec, const void* qkv_op, int batch, int n, int heads, int dh, void* out_op, cudaStream_t st);
// Same contract on the tcgen05 pipe (attention_tc.cu): S and P.V as tcgen05.mma, P kept in TMEM.
int launch_attention_tc(int prec, const void* qkv_op, int batch, int n, int heads, int dh, void* out_op, cudaStream_t st);
bool attention_use_tc();  // capi.cu: true unless PP_ATTENTION=mma is set in the environment

// fp32 token rows (B * hw, c) <-> fp32 NCHW (B, c, hw); NCHW -> operand rows.
int launch_rows_to_nchw(const float* rows, int batch, int hw, int c, float* nchw, cudaStream_t st);
int launch_nchw_to_operand(int prec, const float* nchw, int batch, int hw, int c, void* rows_op, cudaStream_t st,
                           int pad_gh = 0, int pad_gw = 0);

// ---- head_ops.cu --------------------------------------------------------------------------
// Tap gather ("im2col" of one tap list) between operands:
//   dst[(b, y, x), t * c + ch] = src[(b, y + dy[t], x + dx[t]), c_off + ch]   (0 outside the map)
struct GatherParams {
  int batch, h, w;  // map the rows enumerate
  int c;            // channels taken per tap
  int src_c;        // channels per source row (logical operand width)
  int c_off;        // first source channel
  int ntaps;
  int dy[9], dx[9];
};
int launch_gather_taps(int prec, const GatherParams& p, const void* src_op, void* dst_op, cudaStream_t st);

// fp32 (B, h, w, c) -> MaxPool(ph, pw) -> ReLU -> operand (B * (h / ph) * (w / pw), c)
int launch_pool_relu(int prec, const float* x, int batch, int h, int w, int c, int ph, int pw, void* out_op,
                     cudaStream_t st);

// Tail of the four scalar branches: fp32 (B, 2, 2, 4 * c) -> MaxPool(2, 2) -> ReLU -> per branch
// Conv1x1(c -> k) + bias -> Sigmoid (branches 0..2) / ReLU (branch 3) -> scalars fp32 (B, 4, k).
int launch_branch_tail(const float* x, int batch, int c, int k, const float* w /* (4, k, c) */,
                       const float* bias /* (4, k) */, float* scalars, cudaStream_t st);

// ---- weight packing (engine.cu finalize) ---------------------------------------------------
// ConvTranspose2d(k4, s2, p1) weight (cin, cout, 4, 4) -> phase (py, px) matrix fp32 (cout, 4 * cin),
// column t * cin + ci with t = a * 2 + b, tap a: (dy, ky) = py ? {(+1, 0), (0, 2)} : {(0, 1), (-1, 3)}.
int launch_pack_deconv_phase(const float* w, int cin, int cout, int py, int px, float* out, cudaStream_t st);
// Conv2d 3x3 weight (cout, cin, 3, 3) -> fp32 (cout, 9 * cin), column (ky * 3 + kx) * cin + ci.
int launch_pack_conv3x3(const float* w, int cout, int cin, float* out, cudaStream_t st);
// BatchNorm (eval) folded behind a conv with optional bias: scale = g / sqrt(var + eps),
// shift = (bias - mean) * scale + beta.
int launch_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, const float* conv_bias,
                   float eps, int n, float* scale, float* shift, cudaStream_t st);

// pp_gemm without the C-ABI argument checks (gemm.cu).
int gemm_dispatch(const pp_gemm_args& a, cudaStream_t st);
// `count` (<= 4) GEMMs that differ only in W, a_tap_shift and (up_py, up_px) as one launch (gemm_tc.cu "Grouped launch").
int gemm_dispatch_group(const pp_gemm_args* a, int count, cudaStream_t st);

inline void deconv_tap(int phase, int tap, int* d, int* k) {  // see launch_pack_deconv_phase
  if (phase == 0) { *d = tap == 0 ? 0 : -1; *k = tap == 0 ? 1 : 3; }
  else            { *d = tap == 0 ? 1 : 0;  *k = tap == 0 ? 0 : 2; }
}

}  // namespace pp
