/*
 * egorear_b200 — C-ABI of the B200-native (sm_100a) EgoRear inference hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 * names the reference interface it replaces (paths relative to the EgoRear checkout).
 * The reference itself is pure Python; its only native operator is mmcv's
 * MultiScaleDeformableAttnFunction (pose_estimation/models/utils/deform_attn.py:155-162), so the
 * Python-side binding a maintainer adds is the ctypes stub shown in INTEGRATION.md
 * (egorear_b200/_lib.py is that stub).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the parameter is documented as host;
 *   - the caller (PyTorch) owns every input, output and workspace buffer; the library never
 *     frees or retains them beyond the call, except parameter pointers registered through
 *     egr_*_set_param, which must stay alive until the next egr_*_prepack / destroy;
 *   - derived weights (folded / repacked / bf16 copies) are library-owned inside the handle;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*);
 *     no entry point synchronises the device except create/prepack/destroy;
 *   - return value: 0 on success, otherwise an EGR_ERR_* code; egr_last_error() returns a
 *     thread-local message.  Nothing throws or exits across the ABI;
 *   - there is NO CPU fallback: every compute entry fails with EGR_ERR_NO_DEVICE when the
 *     current device is not an sm_100 (B200) GPU.
 */
#ifndef EGOREAR_B200_H_
#define EGOREAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGR_OK               0
#define EGR_ERR_INVALID      1   /* bad shape / argument (the reference's Python asserts) */
#define EGR_ERR_NO_DEVICE    2   /* no CUDA device, or not sm_100 */
#define EGR_ERR_CUDA         3   /* CUDA runtime / launch error */
#define EGR_ERR_UNSUPPORTED  4   /* configuration outside the shipped configs */
#define EGR_ERR_STATE        5   /* missing parameter / prepack not run / workspace too small */

/* precision of the dense conv/GEMM stages (token/attention math is always fp32) */
#define EGR_PREC_FP32  0   /* fp32 SIMT kernels: reference-grade parity (<=1e-3 rel in fp32) */
#define EGR_PREC_BF16  1   /* bf16 operands, fp32 accumulate in TMEM (tcgen05): stated looser bound 1e-2 (measured 7e-3);
                            * the stages that need more than 8 mantissa bits (pose3d proposal branch, pre-upsample
                            * maps, token Linears) run on fp16 / TF32 operands */
#define EGR_PREC_FP16  3   /* the tensor-core mode inside the fp32 parity bound: every 16-bit stage on fp16 operands
                            * (10-bit mantissa, same bytes and MMA rate as bf16; conversions saturate at +-65504 - the
                            * range is the one assumption), 1x1 / Linear weights as an fp16 hi + lo pair (no weight
                            * rounding error, twice the MMAs on HBM-bound stages), token Linears as "3x TF32"
                            * (x_hi W_hi + x_lo W_hi + x_hi W_lo: fp32-grade) */

const char* egr_last_error(void);
/* process-wide switches.  Plain ints: set them before creating handles and not concurrently with launches ("tc", "wsplit",
 * "tok3x", "asplit", "fold16", "tok_batched", "pose_p2_*" are read when a handle is created / prepacked; "pdl", "pair",
 * "ws", "tail_mma", "stem_fused", "fpn_epi", "conv_prefetch" at every launch):
 *   "tc" (1)            EGR_PREC_BF16 uses the tcgen05 kernels; 0 routes bf16 activations through the SIMT GEMM
 *   "pdl" (1)           programmatic dependent launch between the library's kernels
 *   "tok_batched" (1)   batched token path (token GEMMs on tcgen05) instead of the fused per-frame SIMT kernels
 *   "pose_p2_fp16" (1)  pose3d proposal branch on fp16 operands; 0 = fp32 activations multiplied as TF32
 *   "pose_p2_bf16" (0)  debugging: bf16 operands there (costs the whole 0.1 mm MPJPE budget)
 *   "ws" (0)            weight-stationary mode of the tcgen05 GEMM (measured no gain)
 *   "wsplit" (1)        EGR_PREC_FP16: 0 single fp16 weights, 1 hi + lo pairs for 1x1 / Linear weights, 2 also the 3x3 convs
 *   "tok3x" (1)         EGR_PREC_FP16: mvfex token Linears as "3x TF32" (fp32-grade); 0 = plain TF32 as in EGR_PREC_BF16
 *   "pair" (0)          split weights as paired ring stages [A | W_hi | W_lo] (tiles <= 128 wide: A crosses L2 -> smem once)
 *                       instead of a second k pass over A; -0.4 % step time, other rounding order
 *   "stem_fused" (1)    backbone stem as the fused tcgen05 kernel; 0 = im2col buffer + plain GEMM
 *   "fpn_epi" (0)       backbone FPN upsample-add in the fuse conv's epilogue (measured slower)
 *   "conv_prefetch" (0) 3x3 s1 convs prefetch the next tile's rows into L2 (measured slower)
 *   "asplit" (1)        EGR_PREC_FP16: 32x32 activations of the refine path are written as fp16 pairs [hi | lo] and read as
 *                       split operands: 1 = the F1c, R1a, H2a, H2b outputs (four fewer 10-bit roundings on the way to the
 *                       refined heatmap, ~5 % of the step), 2 = also the 512-channel F1b output (+4 %), 0 = none
 *   "fold16" (1)        EGR_PREC_FP16: the folded memory-projection GEMM of the mvfex tokens on fp16 pairs ([x_hi | x_lo] x
 *                       [W_hi | W_hi | W_lo]) instead of 3x TF32: same three-term product at half the bytes
 *   "tail_mma" (1)      heatmap-head tails with the bilinear upsample on the tensor cores; 0 = CUDA-core interpolation */
int         egr_set_option(const char* key, int value);
int         egr_version(void);
/* stage profiler for bench.py: while enabled the engines record CUDA events between their stages on the launch
 * stream; egr_profile_read aggregates them as "stage:total_ms:count;" into buf (synchronises) and clears them */
int         egr_profile_enable(int on);
int         egr_profile_read(char* buf, int cap);
/* 0 when the current CUDA device is an sm_100 GPU; fills *sm_count / *cc when non-NULL */
int         egr_device_check(int* cc_major_minor, int* sm_count);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
int64_t     egr_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * G1  generate_target                                   replaces generate_heatmap.py:10-48
 *   joints   [n_maps, J, 2] float64, (x, y) pixels in image space (n_maps = frames * views)
 *   out      [n_maps, J, hs, hs] float32, fully written (zeros outside the Gaussian patch)
 *   patch_host: HOST pointer to the (2*3*sigma+1)^2 float32 Gaussian patch exactly as the caller's
 *               numpy evaluates generate_heatmap.py:33-36, or NULL: the library then uses the
 *               recorded numpy values for sigma == 1 and expf otherwise.
 *   3*sigma must be an integer (the reference's slicing breaks otherwise).
 * ------------------------------------------------------------------------------------------- */
int egr_generate_target(const double* joints, float* out, int64_t n_maps, int J, double image_size,
                        int heatmap_size, double sigma, const float* patch_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * D1s get_max_preds_soft_pytorch          replaces pose_estimation/utils/loss.py:145-177 (soft-argmax decoder)
 *   hm [N, J, H, W] float32 -> preds [N, J, 2] float32 = expectation of (x, y) under softmax over H*W,
 *   maxvals [N, J] float32 (the caller views it as [N, J, 1]).  W must be a multiple of 4.
 * ------------------------------------------------------------------------------------------- */
int egr_decode_soft_argmax(const float* hm, int64_t N, int J, int H, int W, int normalize, float* preds, float* maxvals,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * D1i integrate_tensor_2d                 replaces pose_estimation/utils/util.py:80-109 (integral soft-argmax decoder)
 *   hm [N, J, H, W] float32; v = hm * multiplier; p = softmax(v) over H*W (softmax != 0) or relu(v);
 *   coords [N, J, 2] float32 = (sum_w w * sum_h p, sum_h h * sum_w p), divided by sum p in the relu variant;
 *   hm_out [N, J, H, W] float32 = p (second return value of the reference), or NULL.  W must be a multiple of 4.
 * ------------------------------------------------------------------------------------------- */
int egr_integrate_tensor_2d(const float* hm, int64_t N, int J, int H, int W, int softmax, float multiplier, float* coords,
                            float* hm_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * D1  get_max_preds                       replaces pose_estimation/utils/loss.py:122-142
 *   hm [N, J, H, W] float32 -> preds [N, J, 2] float32 (x, y), maxvals [N, J] float32,
 *   valid [N, J] uint8 (torch.bool storage), idx [N, J] int32 flat argmax (may be NULL).
 *   First index wins on ties; NaN propagates like torch.max.  H*W must be a multiple of 4.
 * ------------------------------------------------------------------------------------------- */
int egr_decode_argmax(const float* hm, int64_t N, int J, int H, int W, float threshold, int normalize,
                      float* preds, float* maxvals, uint8_t* valid, int32_t* idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MSDA  single-level multi-scale deformable attention forward
 *       replaces mmcv.ops.multi_scale_deform_attn.MultiScaleDeformableAttnFunction.forward
 *       (call site pose_estimation/models/utils/deform_attn.py:155-162; mmcv==2.2.0)
 *   value [B, H*W, nh, hd] f32; loc [B, Q, nh, 1, P, 2] f32 (x, y in [0,1]);
 *   aw [B, Q, nh, 1, P] f32 (already softmaxed); out [B, Q, nh*hd] f32.
 *   No im2col_step batch-divisibility restriction.
 * ------------------------------------------------------------------------------------------- */
int egr_msda_forward(const float* value, int B, int H, int W, int nh, int hd, const float* loc,
                     const float* aw, int Q, int P, float* out, void* stream);

/* Training side.  Backward of the op above: grad_value [B, H*W, nh, hd] (zeroed by the call, atomic
 * scatter), grad_loc [B, Q, nh, 1, P, 2], grad_aw [B, Q, nh, 1, P], all f32, from grad_out [B, Q, nh*hd] - the three
 * gradients mmcv's MultiScaleDeformableAttnFunction.backward returns. */
int egr_msda_backward(const float* value, int B, int H, int W, int nh, int hd, const float* loc, const float* aw,
                      int Q, int P, const float* grad_out, float* grad_value, float* grad_loc, float* grad_aw, void* stream);
/* nn.MSELoss(reduction="mean") (pl_wrappers/egoposeformer/heatmap_mvf_ex.py:258-261) and MpjpeLoss
 * (models/utils/pose_metric.py:10-16: mean over joints of ||gt - pred||_2; D = 3 coordinates per joint), forward into a
 * device scalar + backward; workspace = egr_loss_workspace_bytes() device bytes; deterministic reductions. */
int64_t egr_loss_workspace_bytes(void);
int egr_mse_loss_forward(const float* pred, const float* target, int64_t n, float* loss, void* workspace, void* stream);
int egr_mse_loss_backward(const float* pred, const float* target, const float* grad_loss, int64_t n, float* grad_pred, void* stream);
int egr_mpjpe_loss_forward(const float* pred, const float* gt, int64_t n_joints, int D, float* loss, void* workspace, void* stream);
int egr_mpjpe_loss_backward(const float* pred, const float* gt, const float* grad_loss, int64_t n_joints, int D, float* grad_pred,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * P3  fisheye reprojection     replaces EgoPoseFormerPose3D._reproject_3d_to_2d
 *     (estimator/egoposeformer_mvf_ex.py:340-382) + FishEyeCameraCalibratedModel.world2camera_pytorch
 *     (utils/camera_models.py:53-104)
 *   pts3d [B, J, 3] f32 cm, MUTATED IN PLACE for syn rigs exactly like the reference
 *   (after the call it holds the reference's left-behind tensor, e.g. p + (12,0,0) for 4-view syn).
 *   cam_ids[n_cams] host ints: 0 FL, 1 FR, 2 BL, 3 BR.   is_rw: 0 syn offsets, 1 use coord_trans_mat
 *   coord_trans_mat [B, n_cams, 4, 4] f32 (rw only, else NULL).
 *   calib_host: HOST float array [4][16]: per camera {cx, cy, size_h, size_w, n_coef, a0..a10} or NULL
 *   for the built-in Ego4View calibration.
 *   anchors_2d [B, n_cams, J, 2] f32, anchors_valid [B, n_cams, J] uint8.
 * ------------------------------------------------------------------------------------------- */
int egr_reproject_fisheye(float* pts3d, int B, int J, const int* cam_ids_host, int n_cams, int is_rw,
                          const float* coord_trans_mat, const float* calib_host, float* anchors_2d,
                          uint8_t* anchors_valid, void* stream);

/* ---------------------------------------------------------------------------------------------
 * H1'  1x1 heatmap head of EgoPoseFormerHeatmap      replaces estimator/egoposeformer_heatmap.py:34-39
 *   feat [N, C, H, W] f32 NCHW, weight [J, C], bias [J] -> out [N, J, H, W] f32
 * ------------------------------------------------------------------------------------------- */
int egr_heatmap_head_1x1(const float* feat, const float* weight, const float* bias, int64_t N, int C, int HW,
                         int J, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense stage (building block of both engines, exported for parity tests and for callers that want a
 * single fused conv/Linear):  D[M,N] = epi(A[M,K] * W[N,K]^T + bias)   replaces one nn.Conv2d(1x1 | 3x3 s2 p1)
 * / nn.Linear + activation of the reference stacks (e.g. estimator/egoposeformer_heatmap_mvf_ex.py:525-532).
 *   activations are channels-last: A row m = pixel (img, y, x), K = channels; W is [N][K] (3x3: [N][ky][kx][Cin]).
 *   amode 0: plain rows (lda); with kblk > 0 the K axis is split in blocks of kblk elements kblk_stride apart.
 *   amode 1: implicit 3x3 stride-2 pad-1 conv over [img][Hin][Win][Cin]; M = n_img*(Hin/2)*(Win/2), K = 9*Cin
 *            (use_tc: K = Cin selects the centre tap only = a 1x1 stride-2 conv, ResNet's downsample branch).
 *   amode 2: implicit 3x3 stride-1 pad-1 conv (use_tc only); M = n_img*Hin*Win, K = 9*Cin  (models/backbones/resnet.py).
 *   epi 0 none, 1 ReLU, 2 exact-erf GELU, 3 ReLU then + relu(bilinear_x2_align_corners(aux)) with
 *       aux [img][(Hout/2)*(Wout/2)][N] in D's dtype (the "offset_pred + frame_feat" of :715);
 *       use_tc, 16-bit D only: 4 relu(acc + bias + aux[M][N]) (BasicBlock residual), 5 relu(acc + bias +
 *       bilinear_x2_align_corners(aux)) (EfficientFPN fuse conv with its low-resolution half commuted in front of the upsample).
 *   groups independent problems at element strides a_gs / w_gs / b_gs / d_gs / aux_gs.
 *   use_tc 1: tcgen05 + TMA kernel, fp32 accumulate in TMEM; a_is_bf16 1: A and W bf16, 2: A and W fp16 (kind::f16),
 *             a_is_bf16 0: A and W fp32, multiplied as TF32 (kind::tf32); d_is_bf16 0 fp32, else the operands' type;
 *   use_tc 0: fp32 SIMT kernel, W fp32, A fp32 or bf16.   d_is_bf16 selects the output dtype.
 * ------------------------------------------------------------------------------------------- */
typedef struct egr_dense_desc {
    const void* A; const void* W; const float* bias; void* D; const void* aux;
    int32_t M, N, K;
    int64_t lda, ldd;
    int32_t amode, epi;
    int32_t kblk; int64_t kblk_stride;
    int32_t Hin, Win, Cin, Hout, Wout;
    int32_t groups;
    int64_t a_gs, w_gs, b_gs, d_gs, aux_gs;
    int32_t a_is_bf16, d_is_bf16, use_tc;
    /* split operands (use_tc only; 0 = plain): A has `ka` columns and the k loop re-reads it from column 0 for k >= ka
     * (ka < K <= 2*ka).  K = 2*ka with W = [W_hi | W_lo] multiplies A by the weight AND its rounding residual (no weight
     * rounding error); K = 3*k0, ka = 2*k0, A = [x | x_lo], W = [W_hi | W_hi | W_lo] is the fp32-grade "3x TF32" product */
    int32_t ka;
    /* pair output (use_tc, 16-bit D): the row of D is [hi (N) | lo (N)], lo = the rounding residual of hi in D's type;
     * ldd >= 2N.  Read by the next stage as the split operand A = [x_hi | x_lo] (its ka = 2N, K = 3N with
     * W = [W_hi | W_hi | W_lo]): no activation rounding between the two stages */
    int32_t out_pair;
} egr_dense_desc;
int egr_dense_stage(const egr_dense_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Heatmap-head tail as a stage (what the mvfex engine runs after H1d / H2c; exposed for the stage-level parity tests):
 *   `nn.Upsample(2, bilinear, align_corners=True), ReLU, Conv2d(128, J, 1)`     egoposeformer_heatmap_mvf_ex.py:108-110, :579-583
 *   z     [G][B][32*32][128] fp16, channels-last pre-activations; w [n_sets][J][128] fp32, bias [n_sets][J] fp32,
 *         group g uses weight set wsel[g] (G <= 4, J <= 16)
 *   hm    fp32, element (b, g, j, y, x) at b*hm_bs + g*hm_gs + (j*64 + y)*64 + x
 *   hm_t  optional 16-bit copy [G][B][J][64*64]: bf16 (precise 0) or fp16 (precise 1)
 *   precise 1: 1x1 weights as an fp16 hi + lo pair, fp32 interpolation (EGR_PREC_FP16); impl 1: tensor-core interpolation
 *   (head_tail_mma.cu), 0: CUDA-core interpolation (head_tail_tc.cu)
 * ------------------------------------------------------------------------------------------- */
int egr_head_tail_stage(const void* z, const float* w, const float* bias, const int32_t* wsel, int B, int G, int J,
                        float* hm, int64_t hm_bs, int64_t hm_gs, void* hm_t, int precise, int impl, void* stream);
/* The refined-feature tail as a stage: out = relu(bilinear_x2_align_corners(z)) as a 16-bit channels-last map
 *   (`nn.Upsample(2, bilinear, align_corners=True), ReLU` of frame_feat_refined_proj_layers, egoposeformer_heatmap_mvf_ex.py:553-560,
 *   with the following 1x1 conv commuted in front of the upsample).
 *   z [n_img][32*32][128] fp16; out [n_img][64*64][128] fp16 (out_f16 1) or bf16 (0); impl 1: tensor-core interpolation
 *   (head_tail_mma.cu), 0: CUDA-core kernel (fp32 interpolation when out_f16) */
int egr_up2_relu_stage(const void* z, int n_img, void* out, int out_f16, int impl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * mvfex engine: everything EgoPoseFormerHeatmapMVFEX.forward does after the backbones
 *   (estimator/egoposeformer_heatmap_mvf_ex.py:262-437, use_1by1_conv=False / jqa configs):
 *   H1 init heads, D1 anchors, and per view HeatmapMVF.forward (:652-731) = Q1 M1 F1 A1 A2 A3 T1 R1 H2.
 *
 *   egr_mvfex_create:   num_views 2 or 4, embed 256 / 4 heads / ffn 512 / 1 layer / 16 points as in
 *                       configs/ego4view_*_heatmap_mvfex-n1_jqa*.yaml (others: EGR_ERR_UNSUPPORTED).
 *   egr_mvfex_set_param: register a parameter of the module's state_dict by its exact key
 *                       (e.g. "heatmap_refiner_front_left.heatmap_proj.0.weight"), fp32 device ptr.
 *   egr_mvfex_prepack:  builds derived weights (folded Wv*Wp, sampled position table, repacked /
 *                       bf16 conv weights).  Must be re-run after parameters change.  Synchronises.
 *   egr_mvfex_workspace_bytes: scratch the caller must provide for a batch of B frames.
 *   egr_mvfex_forward:
 *     feat  [B, V, 128, 64, 64] f32 (FPN output), bfb [B, V, 512, 8, 8] f32 (stride-32 ResNet map)
 *     heatmap_for_anchor [B, V, 15, 64, 64] f32 or NULL (anchors then come from the init heatmap)
 *     out: hm_init, hm_refined [B, V, 15, 64, 64] f32; feat_refined [B, V, 128, 64, 64] f32;
 *          anchors_2d [B, V, 15, 2] f32; anchors_valid [B, V, 15] uint8 (either may be NULL)
 *   egr_mvfex_refiner_forward: one HeatmapMVF.forward (:652-731) — refiner index r in module order
 *     (front_left, front_right, back_left, back_right):
 *     heatmap [B,15,64,64], frame_feat [B,128,64,64], feat_mv [B,V,128,64,64], anchors_2d [B,V,15,2],
 *     anchors_valid [B,V,15] uint8, bfb [B,512,8,8]  ->  hm_refined [B,15,64,64], feat_refined [B,128,64,64]
 * ------------------------------------------------------------------------------------------- */
typedef struct egr_mvfex egr_mvfex;
int egr_mvfex_create(int num_views, int num_heatmap, float heatmap_threshold, int precision, egr_mvfex** out);
int egr_mvfex_destroy(egr_mvfex* h);
int egr_mvfex_set_param(egr_mvfex* h, const char* key, const float* ptr, int64_t numel);
int egr_mvfex_prepack(egr_mvfex* h, void* stream);
int64_t egr_mvfex_workspace_bytes(egr_mvfex* h, int B);
int egr_mvfex_forward(egr_mvfex* h, int B, const float* feat, const float* bfb, const float* heatmap_for_anchor,
                      float* hm_init, float* hm_refined, float* feat_refined, float* anchors_2d,
                      uint8_t* anchors_valid, void* workspace, int64_t workspace_bytes, void* stream);
int egr_mvfex_refiner_forward(egr_mvfex* h, int r, int B, const float* heatmap, const float* frame_feat,
                              const float* feat_mv, const float* anchors_2d, const uint8_t* anchors_valid,
                              const float* bfb, float* hm_refined, float* feat_refined, void* workspace,
                              int64_t workspace_bytes, void* stream);
/* Chaining (EgoPoseFormerMVFEX.forward, estimator/egoposeformer_mvf_ex.py:50-58: heatmap estimator -> pose3d):
 * with export enabled, egr_mvfex_forward keeps in ITS workspace channels-last copies of its input features (activation
 * dtype), of the refined features (activation dtype) and - modes 1 / 3 - a high-precision copy of the refined features
 * for pose3d's proposal branch: fp32 rounded to TF32 (mode 1) or fp16 (mode 3; same 10-bit mantissa, half the bytes).
 * egr_mvfex_staged returns them (valid until the next forward on this handle / workspace).  Passing them to
 * egr_pose3d_use_staged / egr_pose3d_use_staged_final_f16 lets the NEXT egr_pose3d_forward skip re-staging its NCHW
 * inputs (the hints are consumed by that call).  [V][B][64*64][128] layout.
 * With export enabled, egr_mvfex_forward accepts feat_refined == NULL (the chained EgoPoseFormerMVFEX.forward never
 * returns the refined features) and egr_pose3d_forward accepts NULL for an NCHW input whose staged copy was handed over. */
/* mode 4: like 3, but the fp16 copy is the ONLY channels-last copy of the refined features (the refined heatmap head
 * reads it too; egr_mvfex_staged returns refined_nhwc = NULL) - for a pose3d that samples the init features. */
int egr_mvfex_export_staged(egr_mvfex* h, int mode);   /* 0 off, 1 + TF32 copy, 2 activation-dtype copies only, 3 + fp16 copy, 4 */
/* *act_is_bf16 receives the element type of init_nhwc / refined_nhwc: 0 fp32, 1 bf16, 2 fp16 (EGR_PREC_FP16, where the
 * activation copy of the refined features already is the fp16 copy: refined_nhwc_hp == refined_nhwc in every mode) */
int egr_mvfex_staged(egr_mvfex* h, const void** init_nhwc, const void** refined_nhwc, const void** refined_nhwc_hp,
                     int* act_is_bf16);
/* Input hint for a producer that already emits the layout the kernels read (a channels-last bf16 backbone):
 * feat_vmajor_nhwc_bf16 = [V][B][64*64][128] bf16, view-major.  One-shot: the next egr_mvfex_forward skips its
 * NCHW fp32 -> channels-last staging pass, reads this buffer instead (it must stay valid until that forward and any
 * chained pose3d forward have completed) and accepts feat == NULL.  16-bit precisions only (bf16 values under
 * EGR_PREC_BF16, fp16 values under EGR_PREC_FP16); results are bit-identical to passing the fp32 NCHW tensor whose
 * rounding this buffer holds. */
int egr_mvfex_use_staged_input(egr_mvfex* h, const void* feat_vmajor_nhwc_bf16);
/* test/debug: pointer + byte size of a named intermediate inside the workspace of the last forward
 * ("q1", "xT", "t1", "ff", ...); EGR_ERR_INVALID for unknown names */
int egr_mvfex_debug_buffer(egr_mvfex* h, const char* name, void** ptr, int64_t* bytes);

/* ---------------------------------------------------------------------------------------------
 * pose3d engine: EgoPoseFormerPose3D.forward (estimator/egoposeformer_mvf_ex.py:422-452),
 *   conv-MLP proposal branch (use_mlp_avgpool=False, use_mlp_heatmap=False), P1 P2 P3 P4.
 *   camera_model: 0 ego4view_syn, 1 ego4view_rw, 2 syn_stereo_front, 3 rw_stereo_front,
 *                 4 syn_stereo_back, 5 rw_stereo_back
 *   egr_pose3d_forward:
 *     feats_init / feats_final [B, V, 128, 64, 64] f32; coord_trans_mat [B, V, 4, 4] f32 (rw) or NULL
 *     preds [L+1, B, 16, 3] f32: preds[0] = MLP proposal, preds[1..L] = transformer layers (cm)
 * ------------------------------------------------------------------------------------------- */
typedef struct egr_pose3d egr_pose3d;
int egr_pose3d_create(int num_views, int num_joints, int num_layers, int camera_model, int use_pred_heatmap_init,
                      int precision, const float* calib_host, egr_pose3d** out);
int egr_pose3d_destroy(egr_pose3d* h);
int egr_pose3d_set_param(egr_pose3d* h, const char* key, const float* ptr, int64_t numel);
int egr_pose3d_prepack(egr_pose3d* h, void* stream);
int64_t egr_pose3d_workspace_bytes(egr_pose3d* h, int B);
int egr_pose3d_forward(egr_pose3d* h, int B, const float* feats_init, const float* feats_final,
                       const float* coord_trans_mat, float* preds, void* workspace, int64_t workspace_bytes,
                       void* stream);
/* sampled_nhwc: channels-last copy of the map the transformer samples (feats_init when use_pred_heatmap_init, else
 * feats_final), element type as flagged (sampled_is_bf16: 0 fp32, 1 bf16, 2 fp16); final_nhwc_tf32: channels-last fp32 copy of feats_final rounded to TF32.
 * Either may be NULL.  One-shot: consumed by the next egr_pose3d_forward. */
int egr_pose3d_use_staged(egr_pose3d* h, const void* sampled_nhwc, int sampled_is_bf16, const float* final_nhwc_tf32);
/* Hint for feats_final when the proposal branch runs in fp16 (egr_pose3d_proposal_dtype == 3): its channels-last fp16
 * copy (what egr_mvfex_export_staged(h, 3) leaves).  One-shot. */
int egr_pose3d_use_staged_final_f16(egr_pose3d* h, const void* final_nhwc_f16);
/* ---------------------------------------------------------------------------------------------
 * backbone engine: ResnetBackbone.forward of the two stereo estimators
 *   (pose_estimation/models/backbones/resnet.py:6-152: torchvision ResNet18 + EfficientFPN, eval mode, BN folded) on the
 *   tcgen05 conv stages.  Parameters by state_dict key of EgoPoseFormerHeatmapMVFEX
 *   ("heatmap_estimator_stereo_front.encoder.backbone.layer_s2.0.weight", ... incl. the BN running statistics).
 *   egr_backbone_forward: img [B, V, 3, 256, 256] f32 ->
 *     feat_staged [V, B, 64*64, 128] 16-bit (bf16 under EGR_PREC_BF16, fp16 under EGR_PREC_FP16): exactly what
 *       egr_mvfex_use_staged_input takes - the FPN map never exists in NCHW fp32;
 *     bfb [B, V, 512, 8, 8] f32: the stride-32 ResNet map (`backbone_feat_bottom` of the jqa query).
 *   EGR_PREC_FP32 is refused (EGR_ERR_UNSUPPORTED): the fp32 path keeps the PyTorch backbone.
 * ------------------------------------------------------------------------------------------- */
typedef struct egr_backbone egr_backbone;
int egr_backbone_create(int num_views, int precision, egr_backbone** out);
int egr_backbone_destroy(egr_backbone* h);
int egr_backbone_set_param(egr_backbone* h, const char* key, const float* ptr, int64_t numel);
int egr_backbone_prepack(egr_backbone* h, void* stream);
int64_t egr_backbone_workspace_bytes(egr_backbone* h, int B);
int egr_backbone_forward(egr_backbone* h, int B, const float* img, void* feat_staged, float* bfb, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* operand type of the proposal branch (conv_frame_feat + mlp_pred.0): 0 fp32 SIMT, 1 bf16, 2 TF32, 3 fp16 */
int egr_pose3d_proposal_dtype(egr_pose3d* h);
int egr_pose3d_debug_buffer(egr_pose3d* h, const char* name, void** ptr, int64_t* bytes);

/* ---------------------------------------------------------------------------------------------
 * Result packing for the multi-GPU all-gather: 2D joints of the final heatmap and the
 * final 3D pose into one fp32 row per frame  [B, V*J2*2 + J3*3]  (672 B/frame for 4 views).
 * ------------------------------------------------------------------------------------------- */
int egr_pack_joints(const float* preds2d, const float* pose3d, int B, int n2d, int n3d, float* packed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * E1  eval-time heatmap metrics
 *     replaces pose_estimation/pl_wrappers/egoposeformer/heatmap_mvf_ex.py:263-299 (`evaluate`) and
 *              pose_estimation/pl_wrappers/egoposeformer/pose_3d_mvf_ex.py:335-360 (`evaluate_heatmap`)
 *   pred, gt  [B, V, C, H, W] float32; the [V,C,H,W] block of a frame is contiguous, consecutive frames are
 *             *_batch_stride ELEMENTS apart (so the wrappers' view slices pred[:, 0:2] need no copy).
 *   l1, pos_l1 [B] float32: sum over views/joints/pixels of |pred-gt|, and of the same where gt > 0.
 *   scalars   [2] float32: {MSELoss(pred, gt), MSELoss of the arg-max pixel coordinates of pred and gt, both
 *             multiplied by gt's validity mask (gt max >= threshold; the wrappers use threshold 1.0)}.
 *   workspace: egr_eval_heatmap_workspace_bytes(B, V, C) bytes of device memory, 16-byte aligned.
 *   Every map is read once (2 x H*W*4 bytes per (frame, view, joint)); reductions run in a fixed order.
 * ------------------------------------------------------------------------------------------- */
size_t egr_eval_heatmap_workspace_bytes(int64_t B, int V, int C);
int egr_eval_heatmap(const float* pred, int64_t pred_batch_stride, const float* gt, int64_t gt_batch_stride, int64_t B,
                     int V, int C, int H, int W, float threshold, float* l1, float* pos_l1, float* scalars,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * E2  eval-time pose metrics
 *     replaces pose_estimation/pl_wrappers/egoposeformer/pose_3d_mvf_ex.py:317-333 (`evaluate_pose`):
 *     batch_compute_similarity_transform_numpy (models/utils/pose_metric.py:104-167, a per-sample numpy SVD loop),
 *     compute_mpjpe_batch / compute_pck_3d_batch / compute_auc_3d_batch (utils/loss.py:9-48).
 *   pred, gt [B, J, 3] float32 contiguous, model units (cm); J <= 32.
 *   metrics  [B, 4] float64: {MPJPE * unit_scale, PA-MPJPE * unit_scale, PCK(pck_threshold) * 100,
 *            AUC(auc_thresholds) * 100}; PCK/AUC compare distances of the scaled poses (pred*unit_scale ...).
 *            MPJPE/PCK/AUC are fp32 values widened; PA-MPJPE is float64 like the reference's numpy path.
 *   auc_thresholds: DEVICE pointer to n_auc <= 64 float32 thresholds (np.linspace(0, 150, 31) in the reference).
 *   s1_hat   [B, J, 3] float64 or NULL: the similarity-aligned prediction.
 * ------------------------------------------------------------------------------------------- */
int egr_eval_pose(const float* pred, const float* gt, int64_t B, int J, float unit_scale, float pck_threshold,
                  const float* auc_thresholds, int n_auc, double* metrics, double* s1_hat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * I1  image preprocessing
 *     replaces, per image, `transform(Image.open(p).convert("RGB").resize([256, 256], Image.BICUBIC))` with
 *     transform = Compose([ToTensor(), Normalize(mean, std)])
 *     (pose_estimation/datasets/ego4view_rw/ego4view_rw_heatmap_mvf.py:40-41,96-99 and the five sibling datasets);
 *     bit-exact with Pillow's 8-bit two-pass antialiased bicubic resampler + torchvision's float32 ops.
 *   images     [N, Hin, Win, 3] uint8 decoded RGB (device), N <= 65535
 *   out        [N, 3, Hout, Wout] float32 = ((resized / 255) - mean) / std
 *   resized_u8 [N, Hout, Wout, 3] uint8 = the resized image itself, or NULL
 *   mean3_host, std3_host: HOST pointers to 3 floats each.
 *   Strips with Hin > 100 * Win that shrink vertically return EGR_ERR_UNSUPPORTED: Pillow >= 11 resamples those
 *   vertically first (PIL/Image.py, Image.resize), which changes the 8-bit intermediate image.
 *   The first call for a given (Hin -> Hout, Win -> Wout) on a device allocates and uploads that geometry's coefficient
 *   tables (a few KB, kept for the life of the process): make it outside CUDA-graph capture.
 * ------------------------------------------------------------------------------------------- */
/* host-only helper (no device): Pillow's bicubic coefficient table for in_size -> out_size as the kernel uses it.
 * *ksize = taps per output; bounds [out_size][2] = (first input index, tap count); kk [out_size][ksize] 22-bit fixed
 * point.  bounds / kk may be NULL (query ksize first). */
int egr_resample_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk);
/* host-only: the same coefficients as the kernel's dp4a operands: *ks4 = ceil(ksize / 4) steps per output; digits
 * [out_size][ks4][4] words = {d0, d1, d2, 0}, each packing the signed 8-bit digits of 4 consecutive taps, with
 * k = d0 + 256 * d1 + 65536 * d2.  digits may be NULL (query ks4 first). */
int egr_resample_digits(int in_size, int out_size, int* ks4, uint32_t* digits);
int egr_preprocess_images(const uint8_t* images, int64_t N, int Hin, int Win, int Hout, int Wout, const float* mean3_host,
                          const float* std3_host, float* out, uint8_t* resized_u8, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGOREAR_B200_H_ */
