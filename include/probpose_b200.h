/*
 * probpose_b200.h - C ABI of the B200-native ProbPose inference hot path.
 *
 * Plain pointers and sizes only; no torch / C++ types.  Every function returns 0 on
 * success or a negative pp_status; pp_last_error() gives the message for the calling
 * thread.  Device pointers are raw CUDA device addresses, `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).  Nothing here allocates device
 * memory: the caller sizes one workspace with pp_engine_workspace_bytes() and owns it.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference checkout, MiraPurkrabek/ProbPose_code @ 93bc991).
 */
#ifndef PROBPOSE_B200_H
#define PROBPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PP_API __attribute__((visibility("default")))
#else
#define PP_API
#endif

#define PP_MAX_KEYPOINTS 17 /* OKS sigma table length, mmpose/codecs/utils/post_processing.py:16 */
#define PP_RECORD_FLOATS 7  /* [x_hm, y_hm, conf, prob, vis, oks, err] per keypoint */

typedef enum pp_status {
  PP_OK = 0,
  PP_ERR_INVALID = -1,     /* bad argument (the Python layer raises ValueError)           */
  PP_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                          */
  PP_ERR_UNSUPPORTED = -3, /* shape or device outside what the sm_100a kernels implement   */
  PP_ERR_STATE = -4        /* engine used before its weights were loaded, etc.             */
} pp_status;

/* GEMM arithmetic of the tensor-core path (operands staged by TMA, tcgen05.mma, fp32
 * accumulate in TMEM).  FP16X3 is the parity mode: every operand is split into
 * fp16 hi + scaled fp16 lo and three MMAs reproduce fp32-grade products. */
typedef enum pp_precision {
  PP_PREC_FP16X3 = 0, /* parity mode   (~2^-22 relative operand error)                    */
  PP_PREC_BF16 = 1,   /* throughput mode, bf16 operands                                   */
  PP_PREC_FP16 = 2,   /* throughput mode, fp16 operands                                   */
  PP_PREC_FP32_SIMT = 3 /* CUDA-core fp32 FFMA GEMMs; verification of the tensor path     */
} pp_precision;

PP_API const char* pp_last_error(void);
/* Library / build identification: "probpose_b200 <version> sm_100a". */
PP_API const char* pp_version(void);

/* ------------------------------------------------------------------------------------
 * Fused decode: [sparsemax(logits / T) * normalize -> clamp] x {pass, flipped pass}
 * -> flip-TTA merge -> per-keypoint OKS-Gaussian convolution (reflect border) ->
 * first-max argmax -> quadratic sub-pixel step -> score lookup -> record.
 *
 * Replaces, in one kernel launch for the whole batch:
 *   Sparsemax + clamp        mmpose/models/heads/hybrid_heads/probmap_head.py:641-645
 *   flip_heatmaps + average  mmpose/models/utils/tta.py:35-39, probmap_head.py:757-774
 *   BaseHead.decode loop     mmpose/models/heads/base_head.py:57-77 (D2H copy + python loop)
 *   ProbMap.decode           mmpose/codecs/probmap.py:170-220
 *   get_heatmap_expected_value / _get_subpixel_maximums
 *                            mmpose/codecs/utils/post_processing.py:308-430
 * ---------------------------------------------------------------------------------- */
typedef struct pp_decode_cfg {
  int32_t num_keypoints;   /* K <= PP_MAX_KEYPOINTS                                        */
  int32_t height, width;   /* heatmap H, W (64, 48)                                        */
  int32_t input_is_logits; /* 1: maps are raw final-layer logits, run sparsemax here;
                              0: maps are already-normalised heatmaps                      */
  float temperature;       /* probmap_head.py:135 (0.5); used when input_is_logits        */
  float normalize;         /* config `normalize` (1.0); used when input_is_logits         */
  float error_divisor;     /* sqrt(H^2 + W^2), probmap_head.py:786-787; 0 -> computed     */
} pp_decode_cfg;

/*  maps         device (B, K, H, W) fp32
 *  maps_flip    device (B, K, H, W) fp32 from the h-flipped input, or NULL (no TTA)
 *  flip_indices HOST   int32[K] left/right permutation (metainfo["flip_indices"]); may be
 *               NULL when maps_flip is NULL
 *  scalars      device (B, 4, K) fp32 post-activation prob / vis / oks / err, or NULL
 *  scalars_flip device (B, 4, K) fp32 from the flipped pass, or NULL
 *  records      device (B, K, 7) fp32 out: x, y in HEATMAP pixels exactly as the
 *               reference's `locs` (the caller applies probmap.py:218 in float64), conf =
 *               merged heatmap at the integer peak, then prob, vis, oks, err / divisor
 *  merged_out   device (B, K, H, W) fp32 out, or NULL: the merged normalised heatmaps
 *               (test_cfg["output_heatmaps"], probmap_head.py:800-804)                    */
PP_API int pp_decode(const pp_decode_cfg* cfg, const float* maps, const float* maps_flip,
              const int32_t* flip_indices, const float* scalars, const float* scalars_flip,
              int32_t batch, float* records, float* merged_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Tensor-core GEMM building block (exported for tests and profiling):
 *   D[M, N] = epilogue( A[M, K] . W[N, K]^T )      A, W K-major ("TN"), fp32 accumulate
 * Replaces the cuBLASLt / cuDNN calls behind nn.Linear / Conv2d / ConvTranspose2d on
 * this path.
 * ---------------------------------------------------------------------------------- */
typedef enum pp_act { PP_ACT_NONE = 0, PP_ACT_GELU = 1, PP_ACT_RELU = 2 } pp_act;

/* How the epilogue writes a result row m / column n. */
typedef enum pp_out_kind {
  PP_OUT_F32 = 0,      /* fp32 (M, ldd) row-major                                          */
  PP_OUT_OPERAND = 1,  /* next GEMM's A operand in the engine precision (see pp_operand_*) */
  PP_OUT_PLANES = 2    /* fp32 (M / plane, N, plane): channel-major planes (final 1x1 conv
                          -> (B, K, H*W) logits)                                           */
} pp_out_kind;

typedef struct pp_gemm_args {
  int32_t precision;      /* pp_precision                                                  */
  int32_t m, n, k;        /* logical sizes; k % 64 == 0, n % 8 == 0                        */
  const void* a;          /* device operand (see pp_operand_bytes)                         */
  const void* w;          /* device operand, N rows                                        */
  const float* scale;     /* device fp32[n] or NULL (=1): per-column multiplier (BN fold)  */
  const float* shift;     /* device fp32[n] or NULL (=0): per-column bias                  */
  const float* residual;  /* device fp32 (M, ldd) or NULL: added after activation          */
  int32_t act;            /* pp_act                                                        */
  int32_t out_kind;       /* pp_out_kind                                                   */
  void* d;                /* device output                                                 */
  int32_t ldd;            /* PP_OUT_F32: leading dim (floats); PP_OUT_OPERAND: logical K of
                             the produced operand; PP_OUT_PLANES: ignored                  */
  int32_t plane;          /* PP_OUT_PLANES: rows per plane (H*W)                           */
  /* optional row scatter for the 4 sub-pixel phases of ConvTranspose2d(k4,s2,p1):
   * logical row m = (b*hin + i)*win + j is written to ((b*2hin + 2i+py)*2win + 2j+px).    */
  int32_t up_hin, up_win, up_py, up_px; /* up_hin == 0 disables                            */
  int32_t tile_n;         /* 0 = auto; else 32/64/128/192/256 output-tile width (tuning)   */
  int32_t res_mod;        /* > 0: the residual has res_mod rows and output row r adds row
                             r % res_mod (pos_embed broadcast over the batch); 0 = (M, ldd)  */
  /* Implicit-GEMM convolution ("taps"), no im2col copy.  With a_taps > 1 the A operand has
   * logical width k / a_taps and its rows enumerate a ZERO-PADDED NHWC map; K-group t is that
   * operand read at row offset a_tap_shift[t] = dy * pitch + dx (rows outside the operand read
   * as zero), so A.W^T is a 3x3 convolution / one ConvTranspose2d sub-pixel phase evaluated at
   * every padded position.  in_pad selects the padding layout, drops the border rows and maps
   * the interior rows to output row (b, i, j) (or through the up_* scatter):
   *   1  full border:   (in_h + 2) x (in_w + 2) per image, pixel (i, j) at (i + 1, j + 1), pitch in_w + 2
   *   2  shared border: (in_h + 1) x (in_w + 1) per image, pixel (i, j) at (i + 1, j), pitch in_w + 1;
   *      row 0 of every image block is zero (top border of this image = bottom border of the one
   *      before; past the last block the operand reads as zero) and column in_w is zero (right
   *      border = left border of the next row).  15 % extra rows instead of 31 % on a 16 x 12 map:
   *      what pp_engine uses.
   * out_pad (same values) writes into an output map that itself carries such a border (left
   * untouched: the caller zeroes it once), ready to be the tap operand of the next layer.       */
  int32_t a_taps;         /* 0 / 1 = plain operand                                          */
  int32_t a_tap_shift[9];
  int32_t in_pad, in_h, in_w; /* in_h / in_w also required by up_* when in_pad != 0          */
  int32_t out_pad;
  int32_t cta_pair;       /* 0 = auto; 1 = one CTA per 128-row tile; 2 = CTA pairs (tcgen05
                             cta_group::2, 256-row tiles) - tuning / tests                   */
} pp_gemm_args;

PP_API int pp_gemm(const pp_gemm_args* args, void* stream);

/* Bytes of an (rows x k) GEMM operand in `precision`, and the converter from fp32
 * row-major (device -> device) used for weights and by tests. */
PP_API size_t pp_operand_bytes(int32_t precision, int64_t rows, int64_t k);
PP_API int pp_operand_from_f32(int32_t precision, const float* src, int64_t rows, int64_t k, int64_t ld_src,
                        void* dst, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused UDPHeatmap (DARK-UDP) decode - the codec of the ViTPose td-hm configs.
 * Replaces, per person on the host in the reference:
 *   flip-TTA merge        mmpose/models/heads/heatmap_heads/heatmap_head.py:245-256 + tta.py:35-39
 *   BaseHead.decode loop  mmpose/models/heads/base_head.py:57-77 (D2H of the heatmaps)
 *   UDPHeatmap.decode     mmpose/codecs/udp_heatmap.py:146-196 (heatmap_type "gaussian")
 *   get_heatmap_maximum / gaussian_blur   mmpose/codecs/utils/post_processing.py:178-249
 *   refine_keypoints_dark_udp             mmpose/codecs/utils/refinement.py:102-160
 *  maps / maps_flip   device fp32 (B, K, H, W) heatmaps of the plain / flipped pass (maps_flip NULL: no TTA)
 *  records            device fp32 (B, K, 3) out: x, y in heatmap pixels (float32 like the reference's keypoints
 *                     before udp_heatmap.py:194-195), score = maximum of the merged map.  Maps whose maximum is
 *                     <= 0 give (-1, -1) unrefined (the reference refines those with samples wrapped around from
 *                     the neighbouring keypoint's map).
 *  merged_out         optional device fp32 (B, K, H, W): the merged heatmaps (test_cfg output_heatmaps)
 * ---------------------------------------------------------------------------------- */
typedef struct pp_udp_cfg {
  int32_t num_keypoints;    /* K <= PP_MAX_KEYPOINTS                                        */
  int32_t height, width;    /* heatmap size, 64 x 48                                        */
  int32_t blur_kernel_size; /* 11 (sigma 2) / 17 (sigma 3): udp_heatmap.py:88-90            */
} pp_udp_cfg;

PP_API int pp_decode_udp(const pp_udp_cfg* cfg, const float* maps, const float* maps_flip, const int32_t* flip_indices,
                         int32_t batch, float* records, float* merged_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Heatmap read-back: per-person inverse warp of the heatmaps into the (padded) image and the
 * element-wise max over the persons, in one pass.  Replaces, per person on the host in the reference,
 *   revert_heatmap      mmpose/structures/utils.py:146-175  (cv2.warpAffine of the (H, W, K) float heatmap, INTER_LINEAR)
 *   merge_data_samples  mmpose/structures/utils.py:117       (np.max over the per-person full-image tensors)
 *  heatmaps    device fp32 (P, K, H, W)
 *  warp_mats   device fp64 (P, 2, 3): the matrices revert_heatmap passes to cv2.warpAffine (heatmap -> image,
 *              get_warp_matrix(..., inv=True), transforms.py:362-425), computed by the caller as the reference does
 *  out         device fp32 (K, img_h, img_w)
 *  scratch     device fp64 (P, 10) work space (the inverted matrices and each person's output rectangle)
 * Bit-identical to OpenCV's CV_32F warpAffine (fixed-point coordinates, 1/32-pixel float weights, BORDER_CONSTANT 0).
 * ---------------------------------------------------------------------------------- */
PP_API int pp_revert_heatmaps(const float* heatmaps, const double* warp_mats, int32_t persons, int32_t num_keypoints,
                              int32_t height, int32_t width, float* out, int32_t img_h, int32_t img_w, double* scratch,
                              void* stream);

/* ------------------------------------------------------------------------------------
 * Global multi-head self-attention of one ViT layer on the tensor cores.  Replaces
 * mmpretrain 1.2.0 MultiheadAttention.forward between its qkv and proj Linears:
 *   q, k, v = qkv.reshape(B, N, 3, heads, d_h).permute(2, 0, 3, 1, 4)
 *   x = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, heads * d_h)
 * (external package, called from the config's backbone, see SURVEY.md 8c "Backbone").
 *  qkv_op   operand (batch * tokens, 3 * heads * head_dim) in `precision` (the qkv GEMM's
 *           PP_OUT_OPERAND output: [q | k | v] column blocks, heads contiguous inside each)
 *  out_op   operand (batch * tokens, heads * head_dim) out = the proj GEMM's A operand; 32-byte
 *           aligned (the tcgen05 kernel writes whole 32-byte sectors with 256-bit stores)
 *  impl     0 = default (tcgen05: S and P.V on the 5th-gen tensor cores, P kept in TMEM),
 *           1 = mma.sync kernel, 2 = tcgen05 kernel (tests compare the two)
 * Built for tokens == 192 and head_dim 32 / 64; tensor-core precisions only.  The tcgen05 kernel runs persistent CTAs
 * on a static schedule (no device-global scheduler state): launches may overlap freely on different streams and
 * replay from CUDA graphs.
 * ---------------------------------------------------------------------------------- */
PP_API int pp_attention(int32_t precision, const void* qkv_op, int32_t batch, int32_t tokens, int32_t heads,
                        int32_t head_dim, void* out_op, int32_t impl, void* stream);

/* ------------------------------------------------------------------------------------
 * Crop front-end: one frame + per-person forward affine matrices -> the uint8
 * model inputs.  Replaces, per person on the host in the reference, TopdownAffine's
 *   cv2.warpAffine(img, warp_mat, (w, h), flags=cv2.INTER_LINEAR)
 * (mmpose/datasets/transforms/topdown_transforms.py:126) + PackPoseInputs' HWC -> CHW
 * (mmpose/datasets/transforms/formatting.py), bit for bit (OpenCV's fixed-point bilinear warp,
 * BORDER_CONSTANT 0).  warp_mats are the float32 matrices of get_udp_warp_matrix
 * (mmpose/structures/bbox/transforms.py:315-359), computed by the caller as the reference does.
 *  frame_hwc_bgr  device (frame_h, frame_w, 3) uint8, row pitch frame_row_bytes
 *  warp_mats      device (n, 2, 3) fp32
 *  crops          device (n, 3, out_h, out_w) uint8 BGR out (what pp_engine_infer consumes)
 * ---------------------------------------------------------------------------------- */
PP_API int pp_crop_warp(const uint8_t* frame_hwc_bgr, int32_t frame_h, int32_t frame_w, int64_t frame_row_bytes,
                        const float* warp_mats, int32_t n, uint8_t* crops, int32_t out_h, int32_t out_w, void* stream);

/* ------------------------------------------------------------------------------------
 * Engine: the whole forward (preprocess -> ViT -> ProbMapHead -> fused decode).
 * Replaces TopdownPoseEstimator.predict (mmpose/models/pose_estimators/topdown.py:86-126)
 * = PoseDataPreprocessor.forward (models/data_preprocessors/data_preprocessor.py:79-104)
 * + mmpretrain VisionTransformer.forward (config :56-67) + ProbMapHead.predict
 * (probmap_head.py:715-804).
 * ---------------------------------------------------------------------------------- */
typedef struct pp_engine_cfg {
  int32_t precision;        /* pp_precision                                                */
  int32_t max_batch;        /* crops per call, BEFORE flip doubling                        */
  int32_t img_h, img_w;     /* 256, 192                                                    */
  int32_t patch, patch_pad; /* 16, 2                                                       */
  int32_t embed_dim, depth, heads, ffn_dim; /* 384, 12, 12, 1536 (ViT-S) / 768,12,12,3072  */
  int32_t num_keypoints;    /* 17                                                          */
  int32_t deconv_channels;  /* 256 (two deconv layers); 0 = backbone-only engine           */
  float ln_eps;             /* 1e-6                                                        */
  float bn_eps;             /* 1e-5                                                        */
  float temperature, normalize; /* 0.5, 1.0                                                */
  float mean[3], std[3];    /* RGB order, applied after BGR->RGB                           */
  int32_t head_kind;        /* PP_HEAD_PROBMAP (ProbMapHead + ProbMap codec) or PP_HEAD_HEATMAP
                               (HeatmapHead + UDPHeatmap codec: the ViTPose td-hm configs,
                               heads/heatmap_heads/heatmap_head.py:197-268; pp_engine_head then
                               returns the heatmaps and ignores `scalars`, pp_engine_infer writes
                               (B, K, 3) records = x, y, score as pp_decode_udp does)          */
  int32_t blur_kernel_size; /* PP_HEAD_HEATMAP: UDPHeatmap blur kernel (11)                  */
} pp_engine_cfg;

enum { PP_HEAD_PROBMAP = 0, PP_HEAD_HEATMAP = 1 };

typedef struct pp_engine pp_engine;

PP_API size_t pp_engine_workspace_bytes(const pp_engine_cfg* cfg);
/* `workspace` is device memory of at least pp_engine_workspace_bytes(cfg), 1024-byte
 * aligned, owned by the caller and alive until pp_engine_destroy. */
PP_API int pp_engine_create(const pp_engine_cfg* cfg, void* workspace, size_t workspace_bytes, pp_engine** out);
PP_API void pp_engine_destroy(pp_engine* e);

/* Load one tensor by its MMPose state_dict name ("backbone.layers.3.attn.qkv.weight",
 * "head.deconv_layers.0.weight", ...) from fp32 device memory of
 * `numel` elements.  Unknown names return PP_ERR_INVALID.  Call pp_engine_finalize once
 * all tensors are in; it folds BatchNorm and checks nothing is missing. */
PP_API int pp_engine_load(pp_engine* e, const char* name, const float* data, int64_t numel, void* stream);
PP_API int pp_engine_finalize(pp_engine* e, void* stream);

/* Backbone only: x fp32 (B, 3, H, W) normalised RGB -> feat fp32 (B, C, h, w) NCHW,
 * exactly VisionTransformer.forward's featmap (out_type="featmap"). */
PP_API int pp_engine_backbone(pp_engine* e, const float* x, int32_t batch, float* feat_nchw, void* stream);

/* Head forward on NCHW features (ProbMapHead.forward, probmap_head.py:600-625):
 * heat_logits fp32 (B, K, 4h, 4w) raw final-layer output (sparsemax is fused into
 * pp_decode); scalars fp32 (B, 4, K) post-activation prob / vis / oks / err. */
PP_API int pp_engine_head(pp_engine* e, const float* feat_nchw, int32_t batch, float* heat_logits,
                   float* scalars, void* stream);

/* End to end.  Exactly one of crops_u8_bgr (B,3,H,W uint8 BGR, preprocessing fused) and
 * x_f32 (B,3,H,W fp32 already normalised RGB) is non-NULL.  flip_test != 0 runs the
 * mirrored pass too and merges as the reference does.  records: (B, K, 7) fp32 as in
 * pp_decode.  merged_out may be NULL. */
PP_API int pp_engine_infer(pp_engine* e, const uint8_t* crops_u8_bgr, const float* x_f32, int32_t batch,
                    int32_t flip_test, const int32_t* flip_indices, float* records, float* merged_out,
                    void* stream);

/* Number of kernels the last pp_engine_* call launched (bench.py's gpu_launches). */
PP_API int64_t pp_engine_last_launch_count(const pp_engine* e);

/* CUDA-graph replay inside pp_engine_infer.  Everything between the patch extraction (which reads the
 * caller's crops) and the decode (which writes the caller's records) touches the workspace only, so it
 * is captured once per (batch, flip_test) shape - on the second call of that shape - and replayed with
 * one cudaGraphLaunch on the caller's stream afterwards; results are bit-identical to plain launches.
 * Used for calls of at most `max_images` images (flip_test counts twice); 0 switches it off, a negative
 * value removes the limit.  Default 16 (117 launches per call: 0.76 ms of host time for a single crop, 0.02 ms replayed), or the
 * PP_ENGINE_GRAPH environment variable.  While profiling (below) launches are always plain.
 * The reference has no counterpart: mmengine's test_step issues eager PyTorch ops
 * (mmpose/apis/inference.py:194-196). */
PP_API int pp_engine_set_graph(pp_engine* e, int32_t max_images);
/* Number of graph replays so far (tests assert the graph path actually ran). */
PP_API int64_t pp_engine_graph_replay_count(const pp_engine* e);

/* Per-kernel-class device timing (bench.py's roofline numbers).  Between _begin and _end
 * every kernel a pp_engine_* call launches is bracketed by a CUDA event pair on the caller's
 * stream; _end synchronises that stream and sums the elapsed times per class.  gemm_flops is
 * the algorithmic 2*m*n*k of the GEMM launches (the FP16X3 triple issue is NOT counted). */
typedef enum pp_kernel_class {
  PP_KC_GEMM = 0, PP_KC_ATTENTION = 1, PP_KC_DECODE = 2, PP_KC_OTHER = 3, PP_KC_COUNT = 4
} pp_kernel_class;
typedef struct pp_profile {
  double ms[PP_KC_COUNT];
  int64_t launches[PP_KC_COUNT];
  double gemm_flops;
} pp_profile;
PP_API int pp_engine_profile_begin(pp_engine* e);
PP_API int pp_engine_profile_end(pp_engine* e, pp_profile* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Operand range guard.  fp16 operands clamp at +-65504 - in PP_PREC_FP16X3 that is |activation| > 1023.5 after the
 * 64x operand scale (trained ViTs have massive-activation channels; a clamp would be silently wrong keypoints).
 * Every kernel that writes an operand raises a sticky device flag when it clamps a value.  This call synchronises
 * the CURRENT device, returns in *flagged how many kernel groups clamped since the last clear (0 = none) and,
 * with clear != 0, re-arms the flags.  The Python layer checks it after the first call of an engine and
 * periodically, and raises: switch the model to precision "fp32_simt" (no operand range limit).
 * ---------------------------------------------------------------------------------- */
PP_API int pp_operand_overflow(int32_t clear, int32_t* flagged);

/* ------------------------------------------------------------------------------------
 * Multi-GPU: persons are independent, so every rank runs pp_engine_infer on its
 * contiguous shard of the crops and the ONLY exchange is one all-gather of the decoded records.
 * Replaces mmengine's end-of-epoch `collect_results` (pickled python objects through
 * torch.distributed, reached from tools/test.py:136 -> mmengine Evaluator) for this path.
 *  nccl_comm        the caller's ncclComm_t (torch.distributed: ProcessGroupNCCL._comm_ptr()); the library
 *                   resolves ncclAllGather from the libnccl already loaded in the process - it does not link NCCL
 *  send             device fp32, this rank's records (B_local, K, 7) - pass the buffer pp_engine_infer / pp_decode
 *                   wrote (`records`), no staging copy in between
 *  recv             device fp32 (world * B_local, K, 7), rank order = person order for equal contiguous shards
 *  floats_per_rank  B_local * K * 7
 * One ncclAllGather on `stream`; PP_ERR_UNSUPPORTED when no NCCL library is loaded in the process.
 * ---------------------------------------------------------------------------------- */
PP_API int pp_allgather(void* nccl_comm, const float* send, float* recv, int64_t floats_per_rank, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROBPOSE_B200_H */
