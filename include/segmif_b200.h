/* segmif_b200 -- C ABI of the B200 (sm_100a) kernels behind the SegMiF hot path.
 *
 * The reference (JinyuanLiu-CV/SegMiF) has no FFI layer: its hot path is a set of PyTorch
 * nn.Module.forward() methods whose arithmetic is implicit ATen/cuDNN/cuBLAS calls.  Each entry
 * point below replaces one such implicit library call (or a fused chain of them); the comment on
 * every function names the reference call site(s) it replaces as <file>:<lines> relative to the
 * SegMiF repository root.  The Python mirror of the reference classes (segmif_b200/core/*.py)
 * binds these symbols with ctypes (segmif_b200/_lib.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers + sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *    no allocation, no synchronisation and no global state inside (outputs and workspaces are
 *    caller allocated); every launch goes to the stream passed in (a cudaStream_t cast to void*),
 *    which must be the stream the caller orders its other work on.  (The only process-wide state: one-time
 *    cudaFuncSetAttribute opt-ins to large dynamic shared memory, and the read-once diagnostic switches
 *    SEGMIF_WGRAD_TC / SEGMIF_ATTN_TC listed in INTEGRATION.md section 6.)
 *  - return value: 0 on success, negative SEGMIF_ERR_* otherwise; segmif_last_error() returns a
 *    thread-local message for the last failure.
 *  - "tokens"/NHWC: activations are pixel-major [B, H, W, C] (equivalently [B, N, C] tokens, the
 *    layout MiT uses between blocks).  `ld` arguments are the channel pitch of one pixel in
 *    ELEMENTS, `coff` a channel offset inside that pitch: this is how DRDB's dense concatenation
 *    and the decoder / conv2 concatenations are written in place with no torch.cat.
 *  - dtypes are SEGMIF_F32 / SEGMIF_BF16 storage; all arithmetic accumulates in fp32.
 */
#ifndef SEGMIF_B200_H_
#define SEGMIF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGMIF_ABI_VERSION 1

typedef void* segmif_stream_t; /* cudaStream_t */

enum { SEGMIF_F32 = 0, SEGMIF_BF16 = 1 };
enum { SEGMIF_ACT_NONE = 0, SEGMIF_ACT_RELU = 1, SEGMIF_ACT_PRELU = 2, SEGMIF_ACT_GELU = 3 };
enum {
  SEGMIF_OK = 0,
  SEGMIF_ERR_INVALID = -1,  /* bad argument / unsupported shape */
  SEGMIF_ERR_CUDA = -2,     /* CUDA runtime or launch failure  */
  SEGMIF_ERR_DEVICE = -3    /* not an sm_100 device            */
};

int segmif_abi_version(void);
const char* segmif_last_error(void);
/* Checks that `device` is compute capability 10.x and raises the dynamic shared-memory limits of
 * the kernels.  Must be called once per process per device before any other entry point. */
int segmif_init(int device);

/* ---- K2: LayerNorm over the last dim --------------------------------------------------------
 * replaces nn.LayerNorm call sites: core/mix_transformer.py:152-153 (Block.norm1/2, eps 1e-6),
 * :101 (Attention.norm, 1e-5), :196 (OverlapPatchEmbed.norm, 1e-5), :320,328,336,344 (stage norms).
 * x [rows, C] -> y [rows, C]; gamma/beta fp32 [C]; C % 32 == 0, C <= 1024. */
int segmif_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype,
                         int64_t rows, int C, float eps, segmif_stream_t stream);

/* ---- K1/K3/K4/K9/K10/K11/K12: implicit-GEMM convolution / linear layer on tensor cores --------
 * One entry point for every dense contraction on the path (weights pre-packed [Cout][KH*KW][Cin] bf16):
 *   nn.Linear   core/mix_transformer.py:96,102,112 (q, kv, proj), :47,51 (fc1, fc2);
 *               core/segformer_head.py:23 (linear_c*), :77-80 (linear_fuse 1x1 conv + folded BN + ReLU,
 *               linear_pred);                                    -> KH=KW=1, B=1, H=1, W=rows
 *   nn.Conv2d   core/mix_transformer.py:193 (patch_embed2-4, k3 s2 p1), :100 (Attention.sr, k=s=sr);
 *               core/model_fusion.py:135-151 (DRDB Dcov1-5, k3 pad2 dil2 + ReLU, written in place into the
 *               224-channel growth buffer), :155-156 (DRDB 1x1 + ReLU + residual), :1063-1064 (conv2, conv21
 *               + shared PReLU).
 * dst[m, dst_coff+n] = residual[m, res_coff+n] + act( bias[n] + sum_{tap,c} src[pix(m,tap), src_coff+c] * w[n,tap,c] )
 * with m = (b, oy, ox) row-major over [B, Ho, Wo] and out-of-image taps reading zero.  Cin % 32 == 0. */
typedef struct {
  const void* src;        /* bf16 [B, H, W, ld_src]                                   */
  const void* weight;     /* bf16 [Cout, KH*KW, Cin]                                  */
  const float* bias;      /* fp32 [Cout] or NULL                                      */
  const float* prelu_alpha; /* device scalar, read only when act == SEGMIF_ACT_PRELU  */
  const void* residual;   /* [M, ld_res] or NULL, added after the activation          */
  void* dst;              /* [M, ld_dst]                                              */
  int B, H, W, Cin, ld_src, src_coff;
  int KH, KW, stride, pad, dil, Ho, Wo, Cout;
  int act;
  int res_dtype, ld_res, res_coff;
  int dst_dtype, ld_dst, dst_coff;
  const void* pre_add;    /* bf16 [M, ld_pre] or NULL: added BEFORE the activation (segmif_conv3x3_tc_fwd only) --  */
  int ld_pre, pre_coff;   /* the x0-slab partial pre-activation of a DRDB layer, see segmif_drdb_push_tc_fwd        */
} segmif_conv_params;
int segmif_conv_fwd(const segmif_conv_params* p, segmif_stream_t stream);

/* ---- K3 on tcgen05: linear layer / 1x1 conv with TMA-fed 5th-generation tensor cores -------------------------
 * Same contract as segmif_conv_fwd with KH=KW=1 (the nn.Linear call sites listed above and DRDB's 1x1 conv,
 * core/model_fusion.py:155-156), restricted to N % 32 == 0; K % 8 == 0; 16-byte aligned pitches.
 * dst[m, dst_coff+n] = residual[m, res_coff+n] + act(bias[n] + sum_k src[m, src_coff+k] * weight[n, k]).          */
typedef struct {
  const void* src;          /* bf16 [M, ld_src]            */
  const void* weight;       /* bf16 [N, K] (K contiguous)  */
  const float* bias;        /* fp32 [N] or NULL            */
  const float* prelu_alpha; /* device scalar (PReLU only)  */
  const void* residual;     /* [M, ld_res] or NULL         */
  void* dst;                /* [M, ld_dst]                 */
  int M, N, K, ld_src, src_coff;
  int act;
  int res_dtype, ld_res, res_coff;
  int dst_dtype, ld_dst, dst_coff;
  int weight_kn;            /* 0: weight bf16 [N, K] (nn.Linear's own layout: y = x W^T).  1: weight bf16 [K, N] -- the SAME
                               buffer read as the operand of the data gradient dX = dY W (MN-major tcgen05 B operand), so
                               the backward needs no transposed copy of the weights; requires N % 64 == 0               */
  const float* row_scale;   /* optional fp32 [ceil(M / rows_per_scale)]: dst = act(.) * row_scale[row / rows_per_scale] + residual
                               -- timm DropPath in train mode (core/mix_transformer.py:129,152-153) fused into proj / fc2   */
  int rows_per_scale;
} segmif_linear_params;
int segmif_linear_tc_fwd(const segmif_linear_params* p, segmif_stream_t stream);
/* ---- K10/K11 on tcgen05: 3x3 stride-1 'same' convolution (dilation 1 or 2), Cout in {32, 64}, bf16 out ----------
 * DRDB Dcov1-5 (core/model_fusion.py:135-151), conv2 / conv21 (:1063-1064).  Same parameter block as
 * segmif_conv_fwd (KH=KW=3, stride=1, pad=dil, no residual, bias required); weights resident in shared memory,
 * halo tiles fetched by 4-D TMA, nine taps = nine shifted views of one tile.                                     */
int segmif_conv3x3_tc_fwd(const segmif_conv_params* p, segmif_stream_t stream);
/* ---- K10 in push form: one DRDB growth step (core/model_fusion.py:135-151, the five Dcov layers) ---------------
 * Convolution is linear in its input channels, so the five N=32 layer convolutions are evaluated slab by slab: the
 * halo tile of ONE input slab (x0: 64 channels, or a freshly produced g_j: 32 channels) is multiplied with the
 * 3x3 dil-2 weights of EVERY later layer restricted to that slab (n_out = 32 * #layers, up to 128).  Output group i
 * (32 channels) is  dst = [relu]( partial_in + acc + bias )  -- the group that completes a layer writes g_j into the
 * growth buffer, the others update bf16 partial pre-activations.  weight: bf16 [n_out][9*64] (slab 64, tap-major) or
 * [n_out][10*32] (slab 32, taps 0..8 then a zero tap).  All tensors pixel-major bf16.                              */
typedef struct {
  const float* bias;        /* fp32 [32] or NULL                        */
  const void* partial_in;   /* bf16 pixel-major, or NULL                */
  void* dst;                /* bf16 pixel-major                         */
  int ld_partial_in, coff_partial_in, ld_dst, coff_dst, relu;
} segmif_drdb_push_group;
typedef struct {
  const void* src;          /* bf16 [B, H, W, ld_src] growth buffer     */
  const void* weight;
  int B, H, W, ld_src, slab_offset, slab_width, n_out;
  segmif_drdb_push_group groups[4];
} segmif_drdb_push_params;
int segmif_drdb_push_tc_fwd(const segmif_drdb_push_params* p, segmif_stream_t stream);

/* ---- K10 as a dataflow: one whole DRDB (core/model_fusion.py:134-157) as seven CONCURRENT persistent kernels -------
 * x0 push a / b (segmif_drdb_push_tc_fwd's N = 96 / 64 steps), the four pull layers (segmif_conv3x3_tc_fwd with pre_add)
 * and the 1x1 + ReLU + residual (segmif_linear_tc_fwd) run at the same time on disjoint groups of SMs; a stage's tile waits
 * on per-tile-row completion counters of the stage that produces the rows it reads (csrc/dataflow.cuh), so slabs and
 * partial pre-activations are consumed out of L2 a tile row or two after they were written instead of round-tripping
 * through HBM between launches.  Results are bit-identical to the sequential launches.
 * The call forks `stream` into six internal side streams and joins them again (legal under stream capture); `flags` is
 * a device workspace of segmif_drdb_dataflow_workspace_bytes(B, H) bytes that the call zeroes; its LAST word is set to 1
 * if a dependency wait ever timed out (a bug, never expected).  segmif_drdb_dataflow_prepare(device) creates the side
 * streams ahead of time (call it once outside any capture).  This is the only entry point with process-wide state
 * beyond the shared-memory opt-ins.                                                                                   */
typedef struct {
  void* growth;             /* bf16 [B, H, W, ld]: x0 in channels 0..63 on entry; g1..g5 appended at 64..223            */
  int ld;
  void* partial;            /* bf16 [B, H, W, ld_partial] scratch: P2..P5 at channels 0..127                            */
  int ld_partial;
  const void* w_push_a;     /* bf16 [96][9*64]: layers 1, 2, 3 restricted to the x0 slab, tap-major                      */
  const void* w_push_b;     /* bf16 [64][9*64]: layers 4, 5                                                              */
  const void* w_pull[4];    /* bf16 [32][9][32*(j-1)]: layer j = 2..5 restricted to the g-slabs                          */
  const float* bias[5];     /* Dcov1..5                                                                                  */
  const void* w_1x1;        /* bf16 [64][224]                                                                            */
  const float* bias_1x1;
  void* out;                /* bf16 [B*H*W, ld_out] channels out_coff..out_coff+63                                       */
  int ld_out, out_coff;
  int B, H, W;
  unsigned int* flags;
  int ctas[7];              /* SMs per stage (push a, push b, L2, L3, L4, L5, 1x1); all zero = built-in split            */
} segmif_drdb_dataflow_params;
size_t segmif_drdb_dataflow_workspace_bytes(int B, int H);
int segmif_drdb_dataflow_prepare(int device);
int segmif_drdb_dataflow_fwd(const segmif_drdb_dataflow_params* p, segmif_stream_t stream);

/* ---- K1 (stage 1): 7x7 stride-4 pad-3 patch embedding + LayerNorm --------------------------------
 * replaces core/mix_transformer.py:192-198 for patch_embed1, fused with the input affine of
 * Network3.forward core/model_fusion.py:1083-1085 (x*255 - mean)/std  (pass scale=1, shift=0 otherwise).
 * img fp32 NCHW [B,3,H,W]; w fp32 [147][C0] (tap-major: (c*7+ky)*7+kx); tokens fp32 [B, Ho*Wo, C0]. */
int segmif_patch_embed7_ln_fwd(const float* img, const float* w, const float* bias, const float* gamma,
                               const float* beta, float eps, const float* in_scale3, const float* in_shift3,
                               float* tokens, int B, int H, int W, int C0, segmif_stream_t stream);

/* ---- K5: spatial-reduction attention core  softmax(q k^T * scale) v ------------------------------
 * replaces core/mix_transformer.py:107-111 (scores are never materialised).
 * q bf16 [B, N, ldq] (head h at columns h*D..), k/v bf16 [B, Nk, ldkv], out bf16 [B, N, ldo]; D in {32, 64}. */
int segmif_sr_attention_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo,
                            int B, int heads, int N, int Nk, int D, float scale, segmif_stream_t stream);

/* ---- K6: depthwise 3x3 (pad 1, bias) + GELU(erf) on tokens -----------------------------------------
 * replaces core/mix_transformer.py:381-387 (DWConv) + :49 (act).  x,y bf16 [B,H,W,C]; w fp32 [9][C]. */
int segmif_dwconv3x3_gelu_fwd(const void* x, const float* w9c, const float* bias, void* y, int B, int H, int W,
                              int C, segmif_stream_t stream);

/* ---- K8: bilinear resize, align_corners=False, pixel-major --------------------------------------------
 * replaces F.interpolate at core/mix_transformer.py:364-373, core/segformer_head.py:67-73. */
int segmif_bilinear_nhwc_fwd(const void* src, int src_dtype, int B, int h, int w, int C, int ld_src, void* dst,
                             int dst_dtype, int H, int W, int ld_dst, int dst_coff, segmif_stream_t stream);

/* ---- K8+K19: bilinear upsample of class logits fused with argmax (labels bit-exact, lowest index wins ties)
 * replaces test_segmentation.py:170-175.  logits fp32 [B,h,w,nc] pixel-major -> labels int64 [B,H,W]. */
int segmif_upsample_argmax_fwd(const float* logits, int B, int h, int w, int nc, int64_t* labels, int H, int W,
                               segmif_stream_t stream);

/* ---- validation / post-processing on the device ------------------------------------------
 * segmif_confusion_matrix: conf[t*nc + p] += #{i : truth[i] == t, pred[i] == p}, pairs with a value outside [0, nc) dropped
 *   -- sklearn.metrics.confusion_matrix(label, prediction, labels=[0..nc-1]) accumulated as `conf_total += conf`
 *   (test_segmentation.py:173-177); conf is int64 [nc, nc] on the device, rows = ground truth.
 * segmif_fused_to_uint8: val_performance.py:447-460 -- clamp(rgb, 0, 1) -> uint8(255 x) -> NHWC -> batch-wide min-max
 *   renormalisation -> uint8(255 y), numpy's truncating casts and double-precision division reproduced exactly.
 *   rgb fp32 NCHW [B,3,H,W] -> out uint8 [B,H,W,3]; minmax: 2 uint32 of device workspace (receives the min and max). */
int segmif_confusion_matrix(const int64_t* truth, const int64_t* pred, int64_t n, int num_classes, int64_t* conf,
                            segmif_stream_t stream);
int segmif_fused_to_uint8(const float* rgb, unsigned char* out_nhwc, unsigned int* minmax, int B, int64_t HW,
                          segmif_stream_t stream);

/* ---- layout converters at the module boundary (NCHW fp32 is the reference's interface layout) -------- */
int segmif_nhwc_to_nchw(const void* src, int src_dtype, int ld_src, int src_coff, float* dst, int B, int HW, int C,
                        segmif_stream_t stream);
int segmif_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int ld_dst, int dst_coff, int B, int HW, int C,
                        segmif_stream_t stream);

/* ---- K11 (edge layers): 3x3 pad-1 convs with 1 input or 1 output channel + shared PReLU --------------
 * conv1_ir / conv1_vis  core/model_fusion.py:1051-1052,1055-1056: plane fp32 [B,H,W] (channel 0 of an NCHW
 *   tensor with batch stride `bstride` elements) -> bf16 [B,H,W,ld_dst] channels coff..coff+63; w fp32 [9][64].
 * conv22               core/model_fusion.py:1065: bf16 [B,H,W,32] -> fp32 [B,1,H,W]; w fp32 [9][32]. */
int segmif_conv3x3_in1_fwd(const float* plane, int64_t bstride, const float* w, const float* bias,
                           const float* prelu_alpha, void* dst, int ld_dst, int dst_coff, int B, int H, int W,
                           int Cout, segmif_stream_t stream);
int segmif_conv3x3_out1_fwd(const void* src, int ld_src, const float* w, const float* bias, const float* prelu_alpha,
                            float* dst, int B, int H, int W, int Cin, segmif_stream_t stream);

/* ---- K13: hierarchical interactive attention (FeatureFusionModule / CrossPath) -----------------------
 * replaces core/model_fusion.py:453-463 -> :350-361 -> :263-288 (MoAM) and :303-328 (SoAM), two passes:
 *  (1) gram:  G_s = sum_pixels p_s^T p_s  for p_1 = y1, p_2 = y2 (first halves of relu(channel_proj1/2)),
 *             p_3 = u3 (second half of relu(channel_proj3 o conv3|conv4)); because the kv Linears have no
 *             bias, k^T v = Wk G Wv^T, so the 8x8 per-head contexts need only these 64x64 Gram matrices.
 *             partials fp32 [B, nchunk, 3, 64, 64] are reduced deterministically by (2).
 *  (2) ctx:   softmax_{dim=-2}(k^T v * 8^-1/2) per head, folded with end_proj into four 64x64 bf16 matrices
 *             per batch item:  folded [B, 4, 64(out), 64(in)] = {Mz1, Mv1, Mz2, Mv2}.
 *  (3) apply: out_i = LayerNorm(x_i + y3 Mz_i + u_i Mv_i + b_end_i), written pixel-major (ld/coff).
 * x1,x2: bf16 [B,HW,ld] 64 channels; x3: bf16 [B,HW,ld3] with C3 in {64,128} channels.
 * wproj: bf16 packed by the host: [w1y 64x64][w2y 64x64][w3u 64xC3] for gram and
 *        [w3y 64xC3][w1u 64x64][w2u 64x64] for apply (rows = output channel, K-major); bproj fp32 [3][64]. */
int segmif_ffm_gram_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* x3,
                        int ld3, int C3, const void* wproj, const float* bproj, float* partials, int nchunk, int B,
                        int64_t HW, segmif_stream_t stream);
int segmif_ffm_ctx_fwd(const float* partials, int nchunk, const float* wkv /* fp32 [3][128][64]: kv1, kv2, kv3 */,
                       const float* wend /* fp32 [2][64][128] */, void* folded /* bf16 [B,4,64,64] */,
                       float* ctx_out /* fp32 [B,3,8,8,8], required */, int B, segmif_stream_t stream);
int segmif_ffm_apply_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* x3,
                         int ld3, int C3, const void* wproj, const float* bproj, const void* folded,
                         const float* bend /* [2][64] */, const float* ln_gamma /* [2][64] */,
                         const float* ln_beta /* [2][64] */, float eps, void* out1, int ldo1, int coffo1, void* out2,
                         int ldo2, int coffo2, int B, int64_t HW, segmif_stream_t stream);

/* K13 with the segmentation stream kept at the encoder's resolution: relu(channel_proj3(conv3|4(upsample(f)))) equals
 * relu(upsample(Q)), Q = (channel_proj3 o conv3|4)(f) + bias computed once at low resolution (all three maps are linear
 * and bilinear weights sum to one).  q3: bf16 [B, qh, qw, 128] ([0,64) = y3 half, [64,128) = u3 half); the kernels
 * interpolate it per pixel (align_corners=False) instead of reading an upsampled feature map.  wproj / bproj hold only
 * the two remaining projections: gram [w1y][w2y] / [b1y][b2y], apply [w1u][w2u] / [b1u][b2u].                          */
int segmif_ffm_gram_lr_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* q3,
                           int qh, int qw, int H, int W, const void* wproj, const float* bproj, float* partials,
                           int nchunk, int B, segmif_stream_t stream);
int segmif_ffm_apply_lr_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* q3,
                            int qh, int qw, int H, int W, const void* wproj, const float* bproj, const void* folded,
                            const float* bend, const float* ln_gamma, const float* ln_beta, float eps, void* out1,
                            int ldo1, int coffo1, void* out2, int ldo2, int coffo2, int B, segmif_stream_t stream);

/* ---- K14: colour transforms (NCHW fp32) ------------------------------------------------------------------
 * replaces core/model_fusion.py:69-92 (RGB2YCrCb), :94-111 (YCrCb2RGB) and the recompose chain
 * train.py:364-366 / test_fusion.py:102-111 (replace Y by the fused image, back to RGB, clamp to [0,1]). */
int segmif_rgb2ycrcb(const float* rgb, float* ycc, int B, int64_t HW, segmif_stream_t stream);
int segmif_ycrcb2rgb(const float* ycc, float* rgb, int B, int64_t HW, segmif_stream_t stream);
int segmif_recompose_rgb(const float* fused_y, const float* vis_rgb, float* rgb_out, int clamp01, int B, int64_t HW,
                         segmif_stream_t stream);

/* ---- K15-K18: fusion losses, forward (fp32 NCHW single-channel planes [B,1,H,W]) --------------------------
 * Each writes per-block partial sums into `workspace` and reduces them (in fp64, fixed order) into `out`.
 * ssim      pytorch_ssim/__init__.py:19-43  out[0] = mean ssim (size_average) or out[b] per image.
 * laploss2  lap_loss.py:112-118             out[0] = 10*(L1_3 + L1_5) + L1_7 against max(res(ir), res(vis)).
 * laploss   lap_loss.py:93-98               out[0] likewise against res(target).
 * entropy   core/Entropy.py:15-56           out[0] = sum over patches and batch of -sum p log p.
 * sobel_l1  core/loss.py:471-475 + :647-650 out[0] = mean |x - y|, out[1] = mean | sobel(x) - sobel(y) |.
 * mse_l1    F.mse_loss / F.l1_loss          out[0] = mean (x-y)^2, out[1] = mean |x - y|.                  */
size_t segmif_loss_workspace_bytes(int B, int H, int W);
int segmif_ssim_fwd(const float* img1, const float* img2, int B, int H, int W, int per_image, float* workspace,
                    float* out, segmif_stream_t stream);
int segmif_laploss2_fwd(const float* inp, const float* ir, const float* vis, int B, int H, int W, float* workspace,
                        float* out, segmif_stream_t stream);
int segmif_laploss_fwd(const float* inp, const float* target, int B, int H, int W, float* workspace, float* out,
                       segmif_stream_t stream);
int segmif_entropy_fwd(const float* img, int B, int H, int W, int patch, float* workspace, float* out,
                       segmif_stream_t stream);
int segmif_sobel_l1_fwd(const float* x, const float* y, int B, int H, int W, float* workspace, float* out,
                        segmif_stream_t stream);
int segmif_mse_l1_fwd(const float* x, const float* y, int64_t n, float* workspace, float* out,
                      segmif_stream_t stream);
/* CE(ignore_index) over bilinearly upsampled logits: core/model_fusion.py:1095-1096 + train.py:156.
 * logits fp32 [B,h,w,nc] pixel-major, labels int64 [B,H,W]; out[0] = mean over non-ignored pixels, out[1] = their
 * number (the normaliser segmif_upsample_ce_bwd needs): `out` holds TWO floats. */
int segmif_upsample_ce_fwd(const float* logits, int B, int h, int w, int nc, const int64_t* labels, int H, int W,
                           int ignore_index, float* workspace, float* out, segmif_stream_t stream);

/* ============================================================================================================
 * Training side (train.py:266-413, train_fusion): gradients of the fusion losses and of Fusion_Network3_ac.
 * The reference obtains all of these from torch.autograd over the implicit ATen/cuDNN backward kernels; each
 * entry point below replaces the autograd node named in its comment.  Gradients of activations are bf16
 * pixel-major slices (ld / coff as in the forward); parameter gradients are ACCUMULATED (+=) in fp32 straight
 * into caller-provided buffers (the .grad tensors or a flat gradient buffer), so a module applied twice (ffm)
 * needs no extra pass.  `gout*` arguments are DEVICE scalars/vectors holding the upstream gradient(s).
 * ============================================================================================================ */

/* ---- loss backward: gradient w.r.t. the FIRST image argument, fp32 [B,1,H,W]; accumulate != 0 adds into dx.
 * mse_l1   F.mse_loss / F.l1_loss backward (core/loss.py:511-517, :464-476); gout2 = {d/d mse, d/d l1}
 * sobel_l1 Fusionloss3 (core/loss.py:464-476 with Sobelxy :634-650); gout2 = {d/d l1, d/d sobel-l1}
 * ssim     pytorch_ssim/__init__.py:19-43 backward; gout = 1 value (size_average) or B values (per_image)
 * laploss2 / laploss  lap_loss.py:112-118 / :93-98 backward;  entropy  core/Entropy.py:15-56 backward           */
int segmif_mse_l1_bwd(const float* x, const float* y, int64_t n, const float* gout2, float* dx, int accumulate,
                      segmif_stream_t stream);
int segmif_sobel_l1_bwd(const float* x, const float* y, int B, int H, int W, const float* gout2, float* dx,
                        int accumulate, segmif_stream_t stream);
int segmif_ssim_bwd(const float* img1, const float* img2, int B, int H, int W, int per_image, const float* gout,
                    float* dimg1, int accumulate, segmif_stream_t stream);
int segmif_laploss2_bwd(const float* inp, const float* ir, const float* vis, int B, int H, int W, const float* gout,
                        float* dinp, int accumulate, segmif_stream_t stream);
int segmif_laploss_bwd(const float* inp, const float* target, int B, int H, int W, const float* gout, float* dinp,
                       int accumulate, segmif_stream_t stream);
int segmif_entropy_bwd(const float* img, int B, int H, int W, int patch, const float* gout, float* dimg, int accumulate,
                       segmif_stream_t stream);

/* ---- activation backward from the layer OUTPUT y (F.relu core/model_fusion.py:135-156; the shared nn.PReLU
 * :1038,1051-1065, slope must be > 0):  dz = dy * f'(y);  dbias[c] += sum_p dz[p][c];  dalpha += sum dy * z [z<=0].
 * dbias / dalpha may be NULL.  C % 8 == 0, C <= 256.  A PReLU slope <= 0 makes sign(y) ambiguous: the kernels then write
 * NaN gradients (loud) instead of wrong ones; ddp.FusionTrainer additionally checks the slope on the host every 50 steps. */
int segmif_act_bwd(const void* y, int ldy, int coffy, const void* dy, int lddy, int coffdy, void* dz, int lddz,
                   int coffdz, int64_t rows, int C, int act, const float* prelu_alpha, float* dbias, float* dalpha,
                   segmif_stream_t stream);
/* conv22's output plane (core/model_fusion.py:1065): out = prelu(z) fp32 [n], dout fp32 -> dz bf16 at channel
 * coffdz of a pixel-major [n, lddz] tensor (the other channels are left untouched).                              */
int segmif_prelu_plane_bwd(const float* out, const float* dout, int64_t n, const float* prelu_alpha, void* dz, int lddz,
                           int coffdz, float* dbias, float* dalpha, segmif_stream_t stream);
/* out[c] += sum_p x[p][coff + c] (bias gradients); out = a + b on bf16 slices (DRDB residual, :156). */
int segmif_colsum(const void* x, int ld, int coff, int64_t rows, int C, float* out, segmif_stream_t stream);
/* Linear layer backward, parameter side (nn.Linear inside core/mix_transformer.py, core/segformer_head.py, core/model_fusion.py):
 * grad[co*s_co + ci*s_ci] += sum_t dy[t][co] x[t][ci] and, when dbias != NULL, dbias[co] += sum_t dy[t][co], from one pass over
 * dy and x on tcgen05 (both operands MN-major).  workspace: nchunk * Cout * Cin floats, nchunk = segmif_wgrad_lin_chunks(...).
 * Only co < co_take, ci < ci_take are written (padded operands); dbias always receives all Cout sums.                        */
int segmif_wgrad_lin_chunks(int64_t P, int Cin, int Cout);
int segmif_wgrad_lin(const void* dy, int ldy, int coffy, const void* x, int ldx, int coffx, int64_t P, int Cin, int Cout,
                     float* workspace, int nchunk, float* grad, int64_t s_co, int64_t s_ci, int co_take, int ci_take,
                     float* dbias, segmif_stream_t stream);
int segmif_add_bf16(const void* a, int lda, int coffa, const void* b, int ldb, int coffb, void* out, int ldo, int coffo,
                    int64_t rows, int C, segmif_stream_t stream);

/* ---- nn.LayerNorm backward (CrossPath.norm1/2 core/model_fusion.py:347-348,360-361; MiT norms later).
 * x = the LayerNorm INPUT [rows, C] dense; dy, dx pixel-major slices.  dgamma / dbeta / dxsum (column sums of dx,
 * = the gradient of a bias added right before the norm) are accumulated; any may be NULL.  C in {64,128,320,512}. */
int segmif_layernorm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, int lddy, int coffdy,
                         const float* gamma, float eps, void* dx, int dx_dtype, int lddx, int coffdx, int64_t rows, int C,
                         float* dgamma, float* dbeta, float* dxsum, int accumulate /* dx += instead of = */,
                         segmif_stream_t stream);

/* ---- weight gradient of nn.Conv2d (3x3, stride 1, 'same', dilation 1|2: DRDB Dcov1-5, conv1/2/21/22) and of
 * nn.Linear / 1x1 conv (taps = 1; pass B = 1, W = 16, H = ceil(P/16)) as one tensor-core contraction over all pixels:
 *   grad[co*s_co + tap*s_tap + ci*s_ci] += sum_p dy[p][coffy+co] * x[p + tap][coffx+ci]      co < co_take, ci < ci_take
 * dy bf16 [P, ldy], x bf16 [B,H,W,ldx]; Cout % 32 == 0, Cin % 8 == 0; workspace: segmif_wgrad_workspace_bytes().   */
size_t segmif_wgrad_workspace_bytes(int nchunk, int Cout, int taps, int Cin);
int segmif_wgrad_chunks(int B, int H, int W, int64_t P, int Cin, int Cout, int taps, int dil);   /* nchunk to call segmif_wgrad with */
int segmif_wgrad(const void* dy, int ldy, int coffy, const void* x, int ldx, int coffx, int B, int H, int W, int64_t P,
                 int Cin, int Cout, int taps, int dil, float* workspace, int nchunk, float* grad, int64_t s_co,
                 int64_t s_tap, int64_t s_ci, int co_take, int ci_take, segmif_stream_t stream);

/* ---- K13 backward (FeatureFusionModule / CrossPath, core/model_fusion.py:350-361, :263-288, :303-328).
 * Training forward = segmif_ffm_gram_fwd + segmif_ffm_ctx_fwd + segmif_ffm_apply_train_fwd, which also writes the
 * LayerNorm inputs pre1 / pre2 (bf16 [B*HW, 64]).  Backward, with dr_i = segmif_layernorm_bwd(pre_i, dout_i):
 *  bwd_gram : fp32 partials [B, nchunk, 4, 64, 64] of dr1^T y3, dr1^T u1, dr2^T y3, dr2^T u2;
 *  bwd_ctx  : per image: dwend [2][64][128] and dwkv [3][128][64] (fp32, accumulated), and mats bf16 [B,7,64,64] =
 *             {S1, S2, S3, Mz1^T, Mv1^T, Mz2^T, Mv2^T} with S_s = dG_s + dG_s^T;
 *  bwd_apply: dP1, dP2, dP3 bf16 [B*HW, 128] = gradients of the channel_proj1/2/3 pre-activations.
 * wfull bf16 [3][128][64] / bfull fp32 [3][128]: channel_proj1|2|3 weight and bias; x3 has 64 channels.           */
int segmif_ffm_apply_train_fwd(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* x3,
                               int ld3, int C3, const void* wproj, const float* bproj, const void* folded,
                               const float* bend, const float* ln_gamma, const float* ln_beta, float eps, void* out1,
                               int ldo1, int coffo1, void* out2, int ldo2, int coffo2, int B, int64_t HW, void* pre1,
                               void* pre2, segmif_stream_t stream);
int segmif_ffm_bwd_gram(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* x3, int ld3,
                        int coff3, const void* dr1, const void* dr2, const void* wfull, const float* bfull,
                        float* partials, int nchunk, int B, int64_t HW, segmif_stream_t stream);
int segmif_ffm_bwd_ctx(const float* r_partials, int nchunk_r, const float* g_partials, int nchunk_g, const float* ctx,
                       const float* wkv, const float* wend, const void* folded, void* mats, float* dwkv, float* dwend,
                       int B, segmif_stream_t stream);
int segmif_ffm_bwd_apply(const void* x1, int ld1, int coff1, const void* x2, int ld2, int coff2, const void* x3, int ld3,
                         int coff3, const void* dr1, const void* dr2, const void* wfull, const float* bfull,
                         const void* mats, void* dP1, void* dP2, void* dP3, int B, int64_t HW, segmif_stream_t stream);

/* ---- fused AdamW step over a flat fp32 buffer (utils/optimizer.py:16-33 -> torch.optim.AdamW.step):
 * g = grad * grad_scale (1/world_size after the gradient all-reduce); p *= 1 - lr*wd; Adam moments with bias
 * correction for 1-based `step`.                                                                                   */
int segmif_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                      segmif_stream_t stream);

/* ---- segmentation-network training (train.py:115-245 train_seg; the CE term of train_fusion :368) ---------------
 * sr_attention_train_fwd  = segmif_sr_attention_fwd that also stores the per-row log-sum-exp (exp2 domain) [B*heads, N].
 * sr_attention_bwd        core/mix_transformer.py:107-111 backward: dq bf16 (same layout as q); dK / dV ADDED into the
 *                         fp32 accumulator dkv [B, Nk, lddkv] (dK at column h*D, dV at v_off + h*D); D = 64.          */
/* The same contraction on tcgen05 (attention_tc.cu: S = QK^T in TMEM, per-row two-pass softmax, O = PV with V as an MN-major
 * operand) for head dim 64 and Nk <= 320; segmif_sr_attention_fwd / _train_fwd use it when SEGMIF_ATTN_TC=1. lse may be NULL. */
int segmif_sr_attention_tc_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B,
                               int heads, int N, int Nk, int D, float scale, float* lse, segmif_stream_t stream);
/* Flash-style tcgen05 kernel for ANY Nk (attention_fa_tc.cu: 128-key blocks through a TMA ring, two score buffers in TMEM so
 * S_{j+1} = Q K_{j+1}^T runs under the softmax of S_j, online softmax with lazy rescaling of the TMEM accumulator, O += P V with V
 * as an MN-major operand) -- the default of segmif_sr_attention_fwd / _train_fwd for head dim 64 (SEGMIF_ATTN=mma selects the
 * mma.sync kernel); BASELINE configs[3] (MiT-B4 at 1024^2: Nk = 1024).                                                      */
int segmif_sr_attention_fa_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B,
                               int heads, int N, int Nk, int D, float scale, float* lse, segmif_stream_t stream);
int segmif_sr_attention_train_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo,
                                  int B, int heads, int N, int Nk, int D, float scale, float* lse, segmif_stream_t stream);
int segmif_sr_attention_bwd(const void* q, int ldq, const void* k, const void* v, int ldkv, const void* out,
                            const void* dout, int ldo, const float* lse, void* dq, int lddq, float* dkv, int lddkv,
                            int v_off, int B, int heads, int N, int Nk, int D, float scale, segmif_stream_t stream);
/* d logits of mean CrossEntropy(ignore_index) over bilinearly upsampled logits (core/model_fusion.py:1095-1096):
 * gout = upstream gradient (device scalar), count = out[1] of segmif_upsample_ce_fwd; dlogits fp32 [B,h,w,nc].     */
int segmif_upsample_ce_bwd(const float* logits, int B, int h, int w, int nc, const int64_t* labels, int H, int W,
                           int ignore_index, const float* gout, const float* count, float* dlogits,
                           segmif_stream_t stream);
/* adjoint of segmif_bilinear_nhwc_fwd (core/segformer_head.py:67-73): ddst bf16 [B,H,W,ld_dst] slice -> dsrc [B,h,w,C] */
int segmif_bilinear_nhwc_bwd(const void* ddst, int ld_dst, int dst_coff, int H, int W, void* dsrc, int B, int h, int w,
                             int C, segmif_stream_t stream);
/* train-mode nn.BatchNorm2d + ReLU of linear_fuse (core/segformer_head.py:50-55) over z bf16 [rows, C]: batch
 * statistics (biased variance) -> stats fp32 [2][C] = {mean, rstd}; running_mean / running_var updated with `momentum`
 * and the unbiased variance (either may be NULL); workspace = 2*C doubles zeroed by the caller.  Backward: dz, and
 * dgamma / dbeta accumulated.                                                                                        */
int segmif_bn_train_fwd(const void* z, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                        float momentum, float* running_mean, float* running_var, double* workspace, float* stats, void* y,
                        segmif_stream_t stream);
int segmif_bn_train_bwd(const void* z, const void* y, const void* dy, const float* stats, const float* gamma, int64_t rows,
                        int C, double* workspace, void* dz, float* dgamma, float* dbeta, segmif_stream_t stream);
/* The same layer in EVAL mode with gradients (train.py:232-236: val_segformer() leaves the model in eval() and training
 * continues with running statistics, no dropout): stats = {running_mean, rsqrt(running_var + eps)}; dz = gamma * rstd * dy 1[y>0]. */
int segmif_bn_eval_fwd(const void* z, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                       const float* running_mean, const float* running_var, float* stats, void* y, segmif_stream_t stream);
int segmif_bn_eval_bwd(const void* z, const void* y, const void* dy, const float* stats, const float* gamma, int64_t rows,
                       int C, double* workspace, void* dz, float* dgamma, float* dbeta, segmif_stream_t stream);
/* y[b,p,c] = x[b,p,c] * scale[b,c]: nn.Dropout2d forward and backward (core/segformer_head.py:57,79), bf16 [B,HW,C]. */
int segmif_channel_scale(const void* x, const float* scale, void* y, int B, int64_t HW, int C, segmif_stream_t stream);
/* depthwise 3x3 without activation (flip != 0: transposed taps = the data gradient of DWConv, mix_transformer.py:381-387)
 * and the backward of dwconv + GELU w.r.t. the pre-activation: dz = dy * gelu'(dwconv(x) + b), dw9c / dbias accumulated
 * from per-strip partials in `workspace` (fixed summation order: deterministic).                                      */
int segmif_dwconv3x3(const void* x, const float* w9c, const float* bias, void* y, int B, int H, int W, int C, int flip,
                     segmif_stream_t stream);
int64_t segmif_dwconv3x3_gelu_bwd_workspace(int B, int H, int W, int C);      /* number of floats (need not be zeroed) */
int segmif_dwconv3x3_gelu_bwd(const void* x, const float* w9c, const float* bias, const void* dy, void* dz, int B, int H,
                              int W, int C, float* dw9c, float* dbias, float* workspace, segmif_stream_t stream);
/* adjoint of im2col for the strided convolutions (OverlapPatchEmbed.proj mix_transformer.py:193, Attention.sr :100):
 * dcol bf16 [B*Ho*Wo, ldc], column (c*k + ky)*k + kx  ->  dx [B,H,W,C] pixel-major (fp32 or bf16).                   */
int segmif_col2im(const void* dcol, int ldc, void* dx, int dx_dtype, int B, int H, int W, int C, int k, int stride, int pad,
                  segmif_stream_t stream);
/* Network3.forward's x*255 / mean / std as its own op (core/model_fusion.py:1083-1085), the gradient of the colour
 * recompose + clamp (train.py:364-366) w.r.t. the fused Y plane, dtype casts, and x + scale[sample] * y (DropPath). */
int segmif_channel_affine_nchw(const float* x, const float* scale, const float* shift, float* y, int B, int C, int64_t HW,
                               segmif_stream_t stream);
int segmif_recompose_rgb_bwd(const float* rgb, const float* drgb, float* dfused, int clamp01, int B, int64_t HW,
                             segmif_stream_t stream);
int segmif_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t n, segmif_stream_t stream);
int segmif_scale_cast_rows(const float* x, const float* scale, void* y, int64_t rows, int64_t rows_per_sample, int C,
                           segmif_stream_t stream);   /* y = bf16(scale[row / rows_per_sample] * x): DropPath branch gradient */
int segmif_scale_add_rows(const float* x, const void* y, int y_dtype, const float* scale, float* out, int64_t rows,
                          int64_t rows_per_sample, int C, segmif_stream_t stream);

/* ============================================================================================================
 * Strict-precision (fp32-parity) mode.  The reference computes everything in fp32; north_star asks
 * for <= 1e-3 relative end to end and bit-exact argmax labels (test_segmentation.py:169-175).  In this mode activations
 * stay fp32 in HBM; an operand of a tensor-core contraction is its 3-way bf16 split  a = a0 + a1 + a2  stored as three
 * planes, and the contraction evaluates the six partial products of order <= 2^-16 with kind::f16 tcgen05 MMAs into a
 * main and a correction TMEM accumulator (csrc/gemm_split_tc.cu).  The same call sites as above are replaced:
 *   segmif_split_gemm_fwd      nn.Linear / 1x1 conv (row mode: the call sites listed at segmif_linear_tc_fwd) and the
 *                              3x3 / dilated 3x3 nn.Conv2d of core/model_fusion.py:135-151, :1063-1064 (patch mode,
 *                              ntaps = 9; tap_dx / tap_dy = (kx - 1) * dil, (ky - 1) * dil; zero padding by TMA);
 *                              im2col rows from segmif_im2col_split3 serve core/mix_transformer.py:193 (patch_embed2-4)
 *                              and :100 (Attention.sr).
 *   dst[m, n] = residual[m, n] + act(bias[n] + sum_{tap, k} A[pix(m, tap), k] * W[n, tap, k])
 * ============================================================================================================ */
typedef struct {
  const void* a_planes;      /* bf16; plane p starts a_plane_stride ELEMENTS after plane p-1; pixel-major, pitch ld_a     */
  int64_t a_plane_stride;
  const void* w_planes;      /* bf16 [N][ntaps][3][Kp], Kp = 64 * ceil(K / 64), zero padded (segmif_split3 builds it)     */
  const float* bias;         /* fp32 [N] or NULL                                                                          */
  const float* prelu_alpha;  /* device scalar (PReLU only)                                                                */
  const float* residual;     /* fp32 [M, ld_res] or NULL, added after the activation                                      */
  float* dst;                /* fp32 [M, ld_dst] or NULL                                                                  */
  void* dst_planes;          /* bf16 split of the same result, three planes dst_plane_stride apart, or NULL               */
  int64_t dst_plane_stride;
  int M, N, K, ld_a, a_coff;
  int act;                   /* SEGMIF_ACT_NONE | RELU | PRELU                                                            */
  int ld_res, res_coff, ld_dst, dst_coff, ld_dp, dp_coff;
  int nterms;                /* 6: fp32-grade product, 3: two planes (2^-16), 1: plain bf16                               */
  int B, H, W;               /* patch mode (ntaps > 0): A is [B, H, W, ld_a] and M = B*H*W; row mode: ignored             */
  int ntaps;                 /* 0 = row mode                                                                              */
  int tap_dx[9], tap_dy[9];
} segmif_split_gemm_params;
int segmif_split_gemm_fwd(const segmif_split_gemm_params* p, segmif_stream_t stream);
/* x fp32 slice [rows, C] (pitch ld_x, offset coff_x) -> optional ReLU -> y fp32 slice (may alias x, may be NULL) and / or
 * three bf16 planes (planes + p * plane_stride)[r * ld_p + coff_p + c].  Also packs weights: rows = N * ntaps, ld_p = 3 * Kp,
 * plane_stride = Kp over a zero-initialised buffer.                                                                      */
int segmif_split3(const float* x, int ld_x, int coff_x, int64_t rows, int C, int relu, float* y, int ld_y, int coff_y,
                  void* planes, int ld_p, int coff_p, int64_t plane_stride, segmif_stream_t stream);
/* im2col of fp32 pixel-major x [B, H, W, ld_x] (C channels at coff_x) for a k x k / stride / pad convolution, written
 * directly as split planes [3][B*Ho*Wo][k*k*C], column (ky*k + kx)*C + c (core/mix_transformer.py:100,193).              */
int segmif_im2col_split3(const float* x, int ld_x, int coff_x, int B, int H, int W, int C, int k, int stride, int pad,
                         void* planes, int64_t plane_stride, segmif_stream_t stream);
/* core/mix_transformer.py:107-111 on fp32 q [B,N,ldq], k / v [B,Nk,ldkv] -> out fp32 [B,N,ldo]; D in {32, 64}.          */
int segmif_sr_attention_f32_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo,
                                int B, int heads, int N, int Nk, int D, float scale, segmif_stream_t stream);
/* core/mix_transformer.py:381-387 (+ :49 when gelu != 0) on fp32 tokens [B,H,W,C]; w fp32 [9][C].                         */
int segmif_dwconv3x3_f32_fwd(const float* x, const float* w9c, const float* bias, float* y, int B, int H, int W, int C,
                             int gelu, segmif_stream_t stream);
/* K13 pass 1 in strict mode: partials double [B, nchunk, 64, 64] of P^T P over the pixels of each image, P = the fp32
 * 64-channel slice p[.., coff..coff+64) (optionally ReLU'd); pass 2: segmif_ffm_ctx_f64_fwd takes partials
 * [3][B][nchunk][64][64] and produces ctx fp32 [B,3,8,8,8] and the folded matrices fp32 [B,4,64,64] (fp64 inside).       */
int segmif_gram64_f64(const float* p, int ld, int coff, int B, int64_t HW, int relu, double* partials, int nchunk,
                      segmif_stream_t stream);
int segmif_ffm_ctx_f64_fwd(const double* partials, int nchunk, const float* wkv, const float* wend, float* folded,
                           float* ctx_out, int B, segmif_stream_t stream);
/* Generic pieces for the ablation networks (core/model_fusion.py:363-425 CrossPath_M / CrossPath_S at dim 32, :626-1025):
 * xty: partials double [B, nchunk, Cx, Cy] of X^T Y over the pixels of each image (k^T v of CrossAttention[2] :281,316);
 * ctx_blockdiag: ctx fp32 [B, heads, d, d] = softmax_{dim=-2}(scale * k^T v per head) and the block-diagonal [C x C] matrix per
 * image that applies it (W[h d + j][h d + i] = ctx[h][i][j]);  sigmoid_gate: x * sigmoid(x) (AttentionModule :759-771).    */
int segmif_xty_f64(const float* x, int ldx, int coffx, int Cx, const float* y, int ldy, int coffy, int Cy, int B, int64_t HW,
                   double* partials, int nchunk, segmif_stream_t stream);
int segmif_ctx_blockdiag(const double* partials, int nchunk, int C, int heads, float scale, float* ctx, float* wout, int B,
                         segmif_stream_t stream);
int segmif_sigmoid_gate(const float* x, float* out, int64_t n, segmif_stream_t stream);
/* conv1_ir / conv1_vis / conv22 (core/model_fusion.py:1051-1056,1065) with fp32 pixel-major activations.                 */
int segmif_conv3x3_in1_f32_fwd(const float* plane, int64_t bstride, const float* w, const float* bias,
                               const float* prelu_alpha, float* dst, int ld_dst, int dst_coff, int B, int H, int W,
                               int Cout, segmif_stream_t stream);
int segmif_conv3x3_out1_f32_fwd(const float* src, int ld_src, const float* w, const float* bias, const float* prelu_alpha,
                                float* dst, int B, int H, int W, int Cin, segmif_stream_t stream);

/* ---- map-valued loss pieces (core/loss.py:634-650 Sobelxy.forward and the composites that use its MAP: Fusionloss :423-439,
 * Fusionloss4 / Fusionloss_add :545-580, new_loss_sobel :389-399, IQALoss :605-633).  fp32 planes [B,1,H,W].
 * sobel_map: out = |Gx| + |Gy| (zero padding); sobel_map_bwd: dx (+)= d/dx of sum(dout * sobel_map(x)), sign(0) = 0.
 * ew2: mode 0: a*x + b*y;  1: max(x, y);  2: |a + b*x| (y may be NULL);  3: x*y;  4: b * sign(a + b*x) * y.        */
int segmif_sobel_map_fwd(const float* x, float* out, int B, int H, int W, segmif_stream_t stream);
int segmif_sobel_map_bwd(const float* x, const float* dout, float* dx, int B, int H, int W, int accumulate,
                         segmif_stream_t stream);
int segmif_ew2(const float* x, const float* y, float a, float b, int mode, float* out, int64_t n, segmif_stream_t stream);

/* ---- training data path on the device -------------------------------------------------------------
 * Replaces, per batch, what the reference's DataLoader workers do per sample on the host: datasets/voc_fusion3.py:169-209
 * (`VOC12SegDataset.__transforms`) = imutils.random_scaling2 (:34-49 -> _img_rescaling2 :69-91: Pillow BILINEAR on uint8 for
 * the three images, NEAREST for the label), random_fliplr2 (:121-129), PhotoMetricDistortion on the visible image (:295-380:
 * convert, mmcv.bgr2hsv / hsv2bgr = OpenCV), random_crop2 (:199-249: mean_rgb / ignore_index canvas at a random offset, up to
 * ten candidate windows), `/ 255.0` and HWC -> CHW.  Results are bit-identical to the reference's (tests/golden/datapath.npz).
 * The host keeps only the random draws and the accept / reject decision; it describes each sample with this struct
 * (one array in device memory for the kernels, the same array in host memory for argument checks and grid sizing).         */
#define SEGMIF_DP_MAX_OPS 6
#define SEGMIF_DP_OP_CONVERT 0      /* convert(img, alpha, beta): uint8(clip(float32(img) * alpha + beta, 0, 255))  :308-312 */
#define SEGMIF_DP_OP_SATURATION 1   /* hsv[..., 1] = convert(hsv[..., 1], alpha)                                    :332-341 */
#define SEGMIF_DP_OP_HUE 2          /* hsv[..., 0] = (int(hsv[..., 0]) + delta) % 180                                :343-351 */
typedef struct {
  const unsigned char* ir;     /* decoded planes in device memory: infrared [H,W], visible [H,W,3], mask [H,W] or [H,W,3], label [H,W] */
  const unsigned char* vis;
  const unsigned char* mask;
  const unsigned char* label;
  int32_t H, W;                /* decoded size */
  int32_t nh, nw;              /* size after random_scaling2: int(ratio * h), int(ratio * w); == H, W when rescaling is off */
  int32_t resized;             /* 1: the images went through the uint8 resize and are float32 from there on (:75-82)         */
  int32_t flip;                /* random_fliplr2 fired */
  int32_t pad_h, pad_w;        /* H_pad, W_pad of random_crop2 */
  int32_t PH, PW;              /* canvas: max(crop, nh), max(crop, nw) */
  int32_t cand_hs[10], cand_ws[10]; /* the ten candidate windows (H_start, W_start) */
  int32_t hs, ws;              /* the window that was kept (image stage) */
  int32_t n_ops;               /* PhotoMetricDistortion program, in execution order */
  int32_t op_kind[SEGMIF_DP_MAX_OPS];
  int32_t op_u8[SEGMIF_DP_MAX_OPS];   /* the image is uint8 when this op runs (after any convert, or never resized) */
  int32_t op_delta[SEGMIF_DP_MAX_OPS];
  float op_alpha[SEGMIF_DP_MAX_OPS];
  float op_beta[SEGMIF_DP_MAX_OPS];
  int32_t ks_x, ks_y;          /* Pillow ksize per axis: ceil(max(in/out, 1)) * 2 + 1 */
  int32_t mask_c, reserved;    /* mask channels: 1 = plane replicated to three (voc_fusion3.py:46-48), 3 = image (voc_fusion2.py:46) */
  int32_t roi_x0, roi_x1, roi_y0, roi_y1; /* part of the resized image the kept window touches (unflipped coordinates) */
  int32_t src_y0, src_y1;      /* source rows the vertical pass needs for roi_y0..roi_y1 */
  int64_t tab_off;             /* int32 elements into table_arena: 2 nw + nw ks_x + 2 nh + nh ks_y + nw + nh per resized sample */
  /* uint8 intermediates use a row pitch p16(w) = (w + 15) & ~15; offsets are multiples of 16 and the arenas 16-byte aligned */
  int64_t tmp_off;             /* bytes into tmp_arena:     (4 + mask_c) (src_y1 - src_y0) p16(roi_x1 - roi_x0) */
  int64_t rs_off;              /* bytes into resized_arena: (4 + mask_c) (roi_y1 - roi_y0) p16(roi_x1 - roi_x0) */
  int64_t lab_off;             /* bytes into label_arena:   PH p16(PW) */
} segmif_dp_sample;
/* label stage: coefficient / index tables, the label canvas, and for each candidate window stats[n][10][3] = {number of
 * non-ignored values present, count of the most frequent one, non-ignored pixels} (np.unique at :222-225).
 * hist_ws: n * 10 * 256 int32 of device workspace.                                                                         */
int segmif_dp_label_stage(const segmif_dp_sample* samples_dev, const segmif_dp_sample* samples_host, int n, int crop,
                          int ignore_index, int32_t* table_arena, unsigned char* label_arena, int32_t* hist_ws,
                          int32_t* stats, segmif_stream_t stream);
/* image stage: both resize passes over the needed region, then flip + distortion + canvas + crop + /255 + CHW.
 * mean_rgb: 3 floats in HOST memory.  Outputs fp32 [n,3,crop,crop] x3, label fp32 [n,crop,crop] (+ int64 copy if given).   */
int segmif_dp_image_stage(const segmif_dp_sample* samples_dev, const segmif_dp_sample* samples_host, int n, int crop,
                          const float* mean_rgb, const int32_t* table_arena, unsigned char* tmp_arena,
                          unsigned char* resized_arena, const unsigned char* label_arena, float* out_ir, float* out_vis,
                          float* out_mask, float* out_label, int64_t* out_label_i64, segmif_stream_t stream);
/* aug=False branch (:193-205 on uint8 arrays): HWC uint8 (C = 1 is replicated to three channels) -> CHW float64 / 255.0.   */
int segmif_dp_u8_to_chw_f64(const unsigned char* src, int H, int W, int C, double* dst, segmif_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGMIF_B200_H_ */
