// Batched token path (EGR_PREC_BF16 engines): the transformer layers over ALL joint tokens of the batch at once.
//
// The fused per-frame token kernels (token_kernels.cu) are latency-bound: 15/16 tokens per CTA, every CTA re-reading
// every weight from L2.  Here the tokens of the whole batch form one [T = B*J, E] matrix per weight set, every Linear
// runs on the tcgen05 kernel (fp32 tokens multiplied as TF32, operands pre-rounded), and the little that is not a GEMM
// (deformable sampling, 16x16 joint attention, LayerNorm, reprojection) is a handful of small kernels in between.
//
// Cross-attention algebra (reference: MultiViewTransformerLayer._run_cross_attn, egoposeformer_heatmap_mvf_ex.py:874-917;
// EgoPoseFormerTransformerLayer, egoposeformer_mvf_ex.py:540-570; MSDeformAttn.forward, models/utils/deform_attn.py:90-168).
// Everything after the sampling is linear until the residual:
//     fuse( cat_v mask_v( out_proj( value_fold(S_v) + E_v ) ) )
// so it collapses into ONE GEMM   Z[T, E] = A[T, V*KA] * Wc^T + b_fuse   over the sampled row
//     A[t][v] = valid ? [ S(h=0..3)[128 raw ch] | E-term | 1 | 0-pad ] : 0
// with Wc built once at prepack from fuse_mlp, output_proj, value_proj, the memory / feat projection and their biases
// (E-term: the sampled position table P' for mvfex, the in-map bilinear weight mass per head for pose3d).
#pragma once
#include "token_common.cuh"
#include "token_kernels.cuh"

namespace egr {

constexpr int TOK_OA = TOK_NH * TOK_P * 3;   // 192: sampling offsets (128) + attention logits (64) per token

// padded row length of the sampled operand per (token, view)
__host__ __device__ constexpr int tok_ka(int E, bool has_ptab) {
    return ((TOK_NH * TOK_RAWC + (has_ptab ? E : TOK_NH) + 1 + 31) / 32) * 32;      // mvfex 800, pose3d 544
}

// deformable sampling of all (group, frame, joint, view, head); one warp each.
//   oa      [G][T][192] fp32: offsets [h][p][2] then logits [h][p]
//   anchors [B][V][J][2], valid [B][V][J]
//   X       [V][B][H*W][128] channels-last features (float | bf16)
//   ptab    per group [V][H*W][E] bf16 (mvfex) or null;  group g uses ptab[g] (device array of pointers)
//   A       [G][T][V][KA] fp32, rounded to TF32
struct TokSampleArgs {
    int G, B, V, J, H, W, E, KA;
    int split;                          // rows of A are [x | x_lo] (2*V*KA floats), see tok_operand_store
    const float* oa;
    const float* anchors;
    const uint8_t* valid;
    const void* X;
    const __nv_bfloat16* const* ptab;   // device array [G] of bf16 tables, or null
    float* A;
};
int tok_sample(const TokSampleArgs& a, int act_bf16 /*1 bf16, 2 fp16*/, cudaStream_t st);

// Operand layout of the fp32-grade ("3x TF32") token Linears of EGR_PREC_FP16 (gemm.cuh `ka`): a producer called with
// split != 0 writes rows of 2K floats [x | x_lo]: x raw fp32 (the tensor core truncates it to TF32 = x_hi) and
// x_lo = tf32(x - x_hi); with split == 0 it writes one row of K floats rounded to the nearest TF32 value.
#if defined(__CUDACC__)
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float x) { return round_tf32(x - tf32_trunc(x)); }
__device__ __forceinline__ void tok_operand_store(float* row, int c, int lo_off, float v, int split) {
    if (split) { row[c] = v; row[lo_off + c] = tf32_lo(v); }
    else row[c] = round_tf32(v);
}
#endif
// x_lo halves of rows produced by a GEMM epilogue: buf [rows][2K], cols [K, 2K) = tf32_lo(cols [0, K))
int tok_make_lo(float* buf, int64_t rows, int K, cudaStream_t st);

// joint self-attention (SpatialMHA / EgoformerSpatialMHA): qkv [G*B*J][3E] (q | k | v) -> o [G*B*J][E], rounded
int tok_attn(const float* qkv, float* o, int n_frames /*G*B*/, int J, int E, cudaStream_t st, int split = 0);

// out[row] = LayerNorm(res[row] + z[row]) * gamma[g] + beta[g]   (res may be null), rounded to TF32.
//   rows = G * rows_per_group; gamma / beta are device arrays of G pointers
//   split: res / out rows are [x | x_lo] (2E floats; z stays a plain GEMM output of E floats)
//   out2 (optional): out2[row] = LayerNorm(out[row]) * gamma2[g] + beta2[g] in the same launch
int tok_add_ln(const float* res, const float* z, float* out, int G, int rows_per_group, int E, const float* const* gamma,
               const float* const* beta, cudaStream_t st, int split = 0, float* out2 = nullptr,
               const float* const* gamma2 = nullptr, const float* const* beta2 = nullptr);

// mvfex post_norm + token image: LN rows of x [G][B][J][E] -> xT [G][B][E][16] (pos-major, joint-minor, col 15 = 0)
int tok_ln_image(const float* x, void* xT, int xT_bf16, int G, int B, int J, int E, const float* const* gamma,
                 const float* const* beta, cudaStream_t st);

// the same post_norm + token image followed by TransformerHeadLayer's two 1x1 convs (the second commuted in front of the
// upsample) in one tcgen05 kernel (token_head_tc.cu):  x [G][B][J][256] -> t1 [G][B][256 pos][128] bf16
//   w0 [sets][64][16] (15 padded to 16), b0 [sets][64], w3 [sets][128][64], b3 [sets][128]; group g uses set r0 + g
//   ldx: row stride of x in floats (256, or 512 for [x | x_lo] rows); f16: hidden / W3 / t1 in fp16 with W3 as a hi + lo pair
int tok_head_tc(const float* x, int G, int B, int J, const float* const* gamma, const float* const* beta, const float* w0,
                const float* b0, const float* w3, const float* b3, int r0, void* t1, cudaStream_t st, int ldx = 256, int f16 = 0);

// mvfex jqa query input (HeatmapMVF.forward :655-665), in three steps:
//   tok_avgpool:   pooled[g][b][512] = adaptive_avg_pool2d(bfb[g][b], 1), rounded     (bfb of group g, frame b at
//                  bfb + g*bfb_gs + b*bfb_bs: [512][hw])
//   (token GEMM)   vb[g][b][E] = fc_bfb(pooled)
//   tok_add_query: x0[g][b][j][:] = y0[g][b][j][:] + vb[g][b][:] + joint_query_embed[g][j][:], rounded
int tok_avgpool(const float* bfb, int64_t bfb_bs, int64_t bfb_gs, int hw, float* pooled, int G, int B, int C, cudaStream_t st, int split = 0);
int tok_add_query(const float* y0, const float* vb, const float* const* jq, float* x0, int G, int B, int J, int E, cudaStream_t st, int split = 0);

// pose3d: P3 reprojection of the proposal + first query_gen Linear(4 -> E) + ReLU
struct PoseQueryArgs {
    int B, V, J, E, is_rw;
    int cam_id[4];
    CamCalib cam[4];
    const float* ctm;          // [B][V][4][4] or null
    const float* mlp_pred;     // [B][J][3]
    const float* g0_T;         // [4][E]
    const float* g0_b;         // [E]
    float* anchors;            // [B][V][J][2]
    uint8_t* valid;            // [B][V][J]
    float* p3;                 // [B][J][4]: (j+1)/J, x, y, z after the in-place quirk
    float* x0;                 // [B*J][E] rounded
};
int pose_query0(const PoseQueryArgs& a, cudaStream_t st);

// pose3d: preds[b][j][:] = reg_mlp.2(r[b][j]) + p3[b][j][1:4]        r [B*J][E], w2_T [E][3]
int pose_reg_out(const float* r, const float* w2_T, const float* b2, const float* p3, float* preds, int T, int E, cudaStream_t st);

// C[m][n] = sum_k A[m*lda + k] * B[k*ldb + n] (+ bias[n])  — strided, prepack-time only
int small_matmul_ex(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                    int M, int N, int K, cudaStream_t st);

}  // namespace egr
