// mvfex engine: orchestration of the multi-view heatmap refinement hot path
// (EgoPoseFormerHeatmapMVFEX.forward after the backbones, estimator/egoposeformer_heatmap_mvf_ex.py:262-437).
//
// HBM layout.  Activations are channels-last and VIEW-MAJOR ([V][B][H*W][C]) so that every dense stage is a
// row-major [M, K] x [N, K]^T GEMM over a contiguous block of rows per weight set (front/back heads: two views
// each; the four per-view refiners: one view each) and the `groups` of one launch sit at a constant stride.
// Stage dtype is fp32 (EGR_PREC_FP32) or bf16 (EGR_PREC_BF16); module outputs are always fp32 NCHW.
#include "engine_common.cuh"
#include "token_kernels.cuh"
#include "token_batched.cuh"

namespace egr {
int g_opt_tc = 1;
int g_opt_tok_batched = 1;     // EGR_PREC_BF16: batched token path (token GEMMs on tcgen05, TF32) instead of the fused SIMT kernel
int g_opt_wsplit = 1;
int g_opt_tok3x = 1;
int g_opt_asplit = 1;      // option "asplit": EGR_PREC_FP16 keeps four (1) / five (2: also the 512-channel F1b output) 32x32 activations of the refine path as fp16 pairs [hi | lo]
int g_opt_fold16 = 1;      // option "fold16": EGR_PREC_FP16 runs the folded memory-projection GEMM (K = 3200) on fp16 pairs instead of 3x TF32
extern int g_opt_pose_p2_bf16;
extern int g_opt_pose_p2_fp16;
extern int g_opt_ws;
extern int g_opt_stem_fused;
extern int g_opt_fpn_epi;
extern int g_opt_conv_prefetch;
extern int g_opt_pair;
}
using namespace egr;

static const char* kRefiner4[4] = {"heatmap_refiner_front_left", "heatmap_refiner_front_right",
                                   "heatmap_refiner_back_left", "heatmap_refiner_back_right"};
static const char* kHead[2] = {"conv_heatmap_layers_stereo_front", "conv_heatmap_layers_stereo_back"};

constexpr int FH = 64, FW = 64, FHW = 4096, FC = 128;   // feature map geometry of every shipped config
constexpr int EMB = 256, NPOS = 256;

struct egr_mvfex {
    int V = 4, J = 15, prec = EGR_PREC_FP32;
    float thr = 0.5f;
    int head_sets = 2;
    bool packed = false;
    ParamTable params;
    DevPool pool;
    // dense stage weights
    WMat h1_0, h1_2, h1_4, h1_7;              // init heads (sets = head_sets)
    float *h1_9w = nullptr, *h1_9b = nullptr;  // [sets][J][128], [sets][J]
    WMat hp0;                                  // heatmap_proj.0 [256][4096]          (sets = V)
    WMat f1_0, f1_2, f1_4;                     // frame_feat_proj_layers
    WMat t1_0, t1_3;                           // head_layers.0.head.{0,3}  ([64][16] padded, [128][64])
    WMat r1_0, r1_3;                           // frame_feat_refined_proj_layers.0.{0,3}
    WMat h2_0, h2_2, h2_5;                     // conv_heatmap_layers.0.{0,2,5}
    float *h2_7w = nullptr, *h2_7b = nullptr;  // [V][J][128], [V][J]
    MvfTokenW* d_tokw = nullptr;               // device [V]
    // batched token path (bf16 precision): token GEMM weights (fp32 rounded to TF32, sets = V) + per-refiner pointer tables
    bool export_staged = false;                // keep channels-last copies for a chained pose3d forward
    int export_hp = 1;                         // high-precision copy of the refined features: 0 none, 1 fp32/TF32, 2 fp16
    bool refn_f16_only = false;                // the fp16 copy is the only channels-last copy (H2a reads it too)
    __half* h2_0_f16 = nullptr;                // conv_heatmap_layers.0.0 weights [V][256][9*128] in fp16
    const void *st_init = nullptr, *st_refined = nullptr;
    const void* in_staged = nullptr;           // one-shot: the caller's view-major channels-last bf16 copy of the input features
    const void* st_refined_hp = nullptr;
    bool tokb = false;
    int asplit = 0;                            // EGR_PREC_FP16: 1 = F1c / R1a / H2a / H2b write [hi | lo] pairs, their consumers read them (PREC_F16X3); 2 = F1b too
    bool tok3x = false;                        // EGR_PREC_FP16: token operand rows [x | x_lo], weights [W_hi | W_hi | W_lo]
    WMat tk_hp2, tk_fcq, tk_bfb, tk_sa, tk_c, tk_qkv, tk_o, tk_f1, tk_f2;
    const __nv_bfloat16** d_ptab16 = nullptr;  // device [4]: bf16 copies of the sampled position tables
    const float** d_ptrs = nullptr;            // device [TP_COUNT][4]
    int KA = 0;
    bool has_heads = false, has_ref[4] = {false, false, false, false};   // parameter groups present at prepack
    std::unordered_map<std::string, std::pair<void*, int64_t>> dbg;
};

namespace {

enum WKind { W_PLAIN, W_CONV3, W_PAD16 };

// builds one WMat from `sets` parameters "<prefix(set)>.weight/.bias"
template <typename KeyFn>
int make_wmat(egr_mvfex* h, WMat& m, int sets, int N, int K, WKind kind, KeyFn key, const bool* present, cudaStream_t st) {
    m.N = N; m.K = K; m.sets = sets;
    m.f32x3 = nullptr; m.f16x3 = nullptr;      // derived copies of an earlier prepack went with the pool
    if (int rc = h->pool.alloc(&m.f32, (int64_t)sets * N * K)) return rc;
    if (int rc = h->pool.alloc(&m.bias, (int64_t)sets * N)) return rc;
    EGR_CUDA_OK(cudaMemsetAsync(m.f32, 0, sizeof(float) * sets * N * K, st));
    EGR_CUDA_OK(cudaMemsetAsync(m.bias, 0, sizeof(float) * sets * N, st));
    for (int s = 0; s < sets; ++s) {
        if (present && !present[s]) continue;
        int rc = EGR_OK;
        const std::string k = key(s);
        float* dst = m.f32 + (int64_t)s * N * K;
        if (kind == W_PLAIN) {
            const float* w = h->params.get(k + ".weight", (int64_t)N * K, &rc);
            if (!w) return rc;
            EGR_CUDA_OK(cudaMemcpyAsync(dst, w, sizeof(float) * N * K, cudaMemcpyDeviceToDevice, st));
        } else if (kind == W_CONV3) {
            const int Cin = K / 9;
            const float* w = h->params.get(k + ".weight", (int64_t)N * K, &rc);
            if (!w) return rc;
            if ((rc = repack_conv3(w, dst, N, Cin, st))) return rc;
        } else {   // [N][15] -> [N][16], zero padded
            const float* w = h->params.get(k + ".weight", (int64_t)N * (K - 1), &rc);
            if (!w) return rc;
            EGR_CUDA_OK(cudaMemsetAsync(dst, 0, sizeof(float) * N * K, st));
            EGR_CUDA_OK(cudaMemcpy2DAsync(dst, sizeof(float) * K, w, sizeof(float) * (K - 1), sizeof(float) * (K - 1), N,
                                          cudaMemcpyDeviceToDevice, st));
        }
        const float* b = h->params.get(k + ".bias", N, &rc);
        if (!b) return rc;
        EGR_CUDA_OK(cudaMemcpyAsync(m.bias + (int64_t)s * N, b, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
    }
    if (h->prec == EGR_PREC_BF16) {
        if (int rc = h->pool.alloc(&m.bf16, (int64_t)sets * N * K)) return rc;
        if (int rc = cast_bf16(m.f32, m.bf16, (int64_t)sets * N * K, st)) return rc;
    } else if (h->prec == EGR_PREC_FP16 && kind != W_PAD16) {
        // fp16 operands; 1x1 / Linear weights as a hi + lo pair along K (gemm.cuh `ka`), the tensor-bound 3x3 convs single
        const bool split = (kind == W_CONV3) ? g_opt_wsplit >= 2 : g_opt_wsplit >= 1;
        m.kw = split ? 2 : 1;
        if (int rc = h->pool.alloc(&m.f16, (int64_t)sets * N * K * m.kw)) return rc;
        if (split) { if (int rc = split_f16(m.f32, m.f16, (int64_t)sets * N, K, st)) return rc; }
        else if (int rc = cast_f16(m.f32, m.f16, (int64_t)sets * N * K, st)) return rc;
    }
    return EGR_OK;
}

// transposed fp32 copy [K][N] of parameter "<key>.weight" ([N][K])
int make_T(egr_mvfex* h, const std::string& key, int N, int K, const float** out, cudaStream_t st) {
    int rc = EGR_OK;
    const float* w = h->params.get(key + ".weight", (int64_t)N * K, &rc);
    if (!w) return rc;
    float* t = nullptr;
    if ((rc = h->pool.alloc(&t, (int64_t)N * K))) return rc;
    if ((rc = transpose2d(w, t, N, K, st))) return rc;
    *out = t;
    return EGR_OK;
}
int get_vec(egr_mvfex* h, const std::string& key, int n, const float** out) {
    int rc = EGR_OK;
    *out = h->params.get(key, n, &rc);
    return rc;
}

int build_layer(egr_mvfex* h, const std::string& p, const std::string& refiner, LayerW& L, cudaStream_t st) {
    const int E = EMB, V = h->V;
    int rc;
#define T_(field, name, N, K) if ((rc = make_T(h, p + name, N, K, &L.field, st))) return rc
#define V_(field, name, n) if ((rc = get_vec(h, p + name, n, &L.field))) return rc
    T_(so_T, ".cross_attn.sampling_offsets", TOK_NH * TOK_P * 2, E);   V_(so_b, ".cross_attn.sampling_offsets.bias", TOK_NH * TOK_P * 2);
    T_(aw_T, ".cross_attn.attention_weights", TOK_NH * TOK_P, E);      V_(aw_b, ".cross_attn.attention_weights.bias", TOK_NH * TOK_P);
    T_(op_T, ".cross_attn.output_proj", E, E);                         V_(op_b, ".cross_attn.output_proj.bias", E);
    T_(fuse_T, ".fuse_mlp", E, V * E);                                  V_(fuse_b, ".fuse_mlp.bias", E);
    V_(lnc_w, ".norm_cross.weight", E);                                 V_(lnc_b, ".norm_cross.bias", E);
    T_(q_T, ".spatial_attn.q_proj", E, E);                              V_(q_b, ".spatial_attn.q_proj.bias", E);
    T_(k_T, ".spatial_attn.k_proj", E, E);                              V_(k_b, ".spatial_attn.k_proj.bias", E);
    T_(v_T, ".spatial_attn.v_proj", E, E);                              V_(v_b, ".spatial_attn.v_proj.bias", E);
    T_(o_T, ".spatial_attn.out_proj", E, E);                            V_(o_b, ".spatial_attn.out_proj.bias", E);
    V_(lns_w, ".norm_spatial.weight", E);                               V_(lns_b, ".norm_spatial.bias", E);
    T_(f1_T, ".ffn.layers.0.0", TOK_FF, E);                             V_(f1_b, ".ffn.layers.0.0.bias", TOK_FF);
    T_(f2_T, ".ffn.layers.1", E, TOK_FF);                               V_(f2_b, ".ffn.layers.1.bias", E);
    V_(lnf_w, ".norm_ffn.weight", E);                                   V_(lnf_b, ".norm_ffn.bias", E);
#undef T_
#undef V_
    // folded memory projection: mfold_T [128][E] = Wp^T · Wv^T ; ptab [V][HW][E] = (pos + bp) · Wv^T + bv
    const float *Wv, *bv, *Wp, *bp, *pos;
    if ((rc = get_vec(h, p + ".cross_attn.value_proj.weight", E * E, &Wv))) return rc;
    if ((rc = get_vec(h, p + ".cross_attn.value_proj.bias", E, &bv))) return rc;
    if ((rc = get_vec(h, refiner + ".frame_feat_multi_view_proj.weight", E * FC, &Wp))) return rc;
    if ((rc = get_vec(h, refiner + ".frame_feat_multi_view_proj.bias", E, &bp))) return rc;
    if ((rc = get_vec(h, refiner + ".frame_feat_multi_view_pos_embed", (int64_t)V * FHW * E, &pos))) return rc;
    float *WvT, *WpT, *mf, *pp, *pt;
    if ((rc = h->pool.alloc(&WvT, E * E))) return rc;
    if ((rc = h->pool.alloc(&WpT, FC * E))) return rc;
    if ((rc = h->pool.alloc(&mf, FC * E))) return rc;
    if ((rc = h->pool.alloc(&pp, (int64_t)V * FHW * E))) return rc;
    if ((rc = h->pool.alloc(&pt, (int64_t)V * FHW * E))) return rc;
    if ((rc = transpose2d(Wv, WvT, E, E, st))) return rc;
    if ((rc = transpose2d(Wp, WpT, E, FC, st))) return rc;          // [E][128] -> [128][E]
    if ((rc = small_matmul(WpT, WvT, nullptr, mf, FC, E, E, st))) return rc;
    if ((rc = add_rowvec(pos, bp, pp, (int64_t)V * FHW, E, st))) return rc;
    if ((rc = small_matmul(pp, WvT, bv, pt, V * FHW, E, E, st))) return rc;
    L.mfold_T = mf;
    L.ptab = pt;
    L.bfold = nullptr;
    return EGR_OK;
}

struct Bufs {
    void *Xh, *Xown, *h1a, *a1, *b1, *c1, *z, *ff, *r1a, *refn, *hmT, *xT, *h1t, *t1;
    float *q1, *anch, *maxv;
    uint8_t* valid;
    float *tx, *tz, *toa, *tA, *tqkv, *to, *thid, *tpool, *tvb;     // batched token path
    void* refn_hp;                                                  // exported TF32 (fp32) or fp16 channels-last refined features
    float* splitk;                                                  // split-K partial sums of the token GEMMs
};

enum TokPtr { TP_LNC_W, TP_LNC_B, TP_LNS_W, TP_LNS_B, TP_LNF_W, TP_LNF_B, TP_PN_W, TP_PN_B, TP_PTAB, TP_BFB_T, TP_BFB_B, TP_JQ,
              TP_COUNT };

// carve the workspace for B frames and G refiner groups (G = V for the full forward, 1 for one refiner)
int64_t carve(const egr_mvfex* h, int B, int G, bool with_heads, bool own_copy, void* base, int64_t cap, Bufs* o) {
    const int64_t s = is16(h->prec) ? 2 : 4;
    const bool f16m = h->prec == EGR_PREC_FP16;        // the activation copy of the refined features IS the fp16 copy pose3d reads
    const int sp2 = h->tok3x ? 2 : 1;                  // token operand rows [x | x_lo]
    const int V = h->V, J = h->J;
    const int Gh = with_heads ? (V > G ? V : G) : G;   // heads run over all views
    Carver c(base, cap);
    Bufs b{};
    b.Xh = c.take((int64_t)V * B * FHW * FC * s);
    b.Xown = own_copy ? c.take((int64_t)G * B * FHW * FC * s) : nullptr;
    b.h1a = with_heads ? c.take((int64_t)V * B * FHW * FC * s) : nullptr;
    b.a1 = c.take((int64_t)G * B * FHW * 256 * s);
    const int ap = h->asplit ? 2 : 1;                  // pair tensors: rows [hi | lo]
    b.b1 = c.take((int64_t)Gh * B * 1024 * 512 * s * ap);
    b.c1 = c.take((int64_t)Gh * B * 1024 * 256 * s * ap);
    b.z = c.take((int64_t)Gh * B * 1024 * 128 * s);
    b.ff = c.take((int64_t)G * B * 1024 * 128 * s * ap);
    b.r1a = c.take((int64_t)G * B * 1024 * 128 * s * ap);
    b.refn = (!f16m && h->refn_f16_only && G == V) ? nullptr : c.take((int64_t)G * B * FHW * FC * s);
    b.hmT = c.take((int64_t)Gh * B * J * FHW * s);
    b.xT = c.take((int64_t)G * B * NPOS * 16 * s);
    b.h1t = c.take((int64_t)G * B * NPOS * 64 * s);
    b.t1 = c.take((int64_t)G * B * NPOS * 128 * s);
    b.q1 = (float*)c.take((int64_t)G * B * J * EMB * 4 * sp2);
    b.anch = (float*)c.take((int64_t)B * V * J * 2 * 4);
    b.maxv = (float*)c.take((int64_t)B * V * J * 4);
    b.valid = (uint8_t*)c.take((int64_t)B * V * J);
    b.refn_hp = (!f16m && h->export_staged && h->export_hp && G == V) ? c.take((int64_t)G * B * FHW * FC * (h->export_hp == 2 ? 2 : 4)) : nullptr;
    if (h->tokb) {
        const int64_t T = (int64_t)B * J;
        b.tx = (float*)c.take(G * T * EMB * 4 * sp2);
        b.tz = (float*)c.take(G * T * EMB * 4);
        b.toa = (float*)c.take(G * T * TOK_OA * 4);
        b.tA = (float*)c.take(G * T * V * h->KA * 4 * sp2);
        b.tqkv = (float*)c.take(G * T * 3 * EMB * 4);
        b.to = (float*)c.take(G * T * EMB * 4 * sp2);
        b.thid = (float*)c.take(G * T * TOK_FF * 4 * sp2);
        b.tpool = (float*)c.take((int64_t)G * B * 512 * 4 * sp2);
        b.tvb = (float*)c.take((int64_t)G * B * EMB * 4);
    }
    b.splitk = (float*)c.take(SPLITK_SCRATCH_BYTES);
    if (o) *o = b;
    return c.off + 256;
}

// fp32 weight matrix [sets][N][K] + bias [sets][N] without a bf16 copy (TF32 token GEMMs)
int alloc_wmat(egr_mvfex* h, WMat& m, int sets, int N, int K, cudaStream_t st) {
    m.N = N; m.K = K; m.sets = sets; m.bf16 = nullptr;
    if (int rc = h->pool.alloc(&m.f32, (int64_t)sets * N * K)) return rc;
    if (int rc = h->pool.alloc(&m.bias, (int64_t)sets * N)) return rc;
    EGR_CUDA_OK(cudaMemsetAsync(m.f32, 0, sizeof(float) * sets * N * K, st));
    EGR_CUDA_OK(cudaMemsetAsync(m.bias, 0, sizeof(float) * sets * N, st));
    return EGR_OK;
}
int copy_rows(egr_mvfex* h, const std::string& key, float* dst_w, float* dst_b, int N, int K, cudaStream_t st) {
    int rc = EGR_OK;
    const float* w = h->params.get(key + ".weight", (int64_t)N * K, &rc);
    if (!w) return rc;
    const float* b = h->params.get(key + ".bias", N, &rc);
    if (!b) return rc;
    EGR_CUDA_OK(cudaMemcpyAsync(dst_w, w, sizeof(float) * N * K, cudaMemcpyDeviceToDevice, st));
    EGR_CUDA_OK(cudaMemcpyAsync(dst_b, b, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
    return EGR_OK;
}

// derived weights of the batched token path for refiner r (see token_batched.cuh for the algebra)
int build_tokb(egr_mvfex* h, int r, const std::string& refiner, cudaStream_t st) {
    const int E = EMB, V = h->V, KA = h->KA, HD = E / TOK_NH;
    const std::string L = refiner + ".transformer_layers.0";
    int rc;
    if ((rc = copy_rows(h, refiner + ".heatmap_proj.2", h->tk_hp2.f32 + (int64_t)r * E * E, h->tk_hp2.bias + r * E, E, E, st))) return rc;
    if ((rc = copy_rows(h, refiner + ".fc_query.0", h->tk_fcq.f32 + (int64_t)r * E * E, h->tk_fcq.bias + r * E, E, E, st))) return rc;
    if ((rc = copy_rows(h, refiner + ".fc_bfb", h->tk_bfb.f32 + (int64_t)r * E * 512, h->tk_bfb.bias + r * E, E, 512, st))) return rc;
    float* sa = h->tk_sa.f32 + (int64_t)r * TOK_OA * E;
    float* sab = h->tk_sa.bias + r * TOK_OA;
    if ((rc = copy_rows(h, L + ".cross_attn.sampling_offsets", sa, sab, TOK_NH * TOK_P * 2, E, st))) return rc;
    if ((rc = copy_rows(h, L + ".cross_attn.attention_weights", sa + (int64_t)TOK_NH * TOK_P * 2 * E, sab + TOK_NH * TOK_P * 2, TOK_NH * TOK_P, E, st))) return rc;
    float* qkv = h->tk_qkv.f32 + (int64_t)r * 3 * E * E;
    float* qkvb = h->tk_qkv.bias + r * 3 * E;
    if ((rc = copy_rows(h, L + ".spatial_attn.q_proj", qkv, qkvb, E, E, st))) return rc;
    if ((rc = copy_rows(h, L + ".spatial_attn.k_proj", qkv + (int64_t)E * E, qkvb + E, E, E, st))) return rc;
    if ((rc = copy_rows(h, L + ".spatial_attn.v_proj", qkv + (int64_t)2 * E * E, qkvb + 2 * E, E, E, st))) return rc;
    if ((rc = copy_rows(h, L + ".spatial_attn.out_proj", h->tk_o.f32 + (int64_t)r * E * E, h->tk_o.bias + r * E, E, E, st))) return rc;
    if ((rc = copy_rows(h, L + ".ffn.layers.0.0", h->tk_f1.f32 + (int64_t)r * TOK_FF * E, h->tk_f1.bias + r * TOK_FF, TOK_FF, E, st))) return rc;
    if ((rc = copy_rows(h, L + ".ffn.layers.1", h->tk_f2.f32 + (int64_t)r * E * TOK_FF, h->tk_f2.bias + r * E, E, TOK_FF, st))) return rc;
    // cross-attention fold
    const float *Wv, *Wp, *Wop, *bop, *Wfuse, *bfuse;
    if ((rc = get_vec(h, L + ".cross_attn.value_proj.weight", E * E, &Wv))) return rc;
    if ((rc = get_vec(h, refiner + ".frame_feat_multi_view_proj.weight", E * FC, &Wp))) return rc;
    if ((rc = get_vec(h, L + ".cross_attn.output_proj.weight", E * E, &Wop))) return rc;
    if ((rc = get_vec(h, L + ".cross_attn.output_proj.bias", E, &bop))) return rc;
    if ((rc = get_vec(h, L + ".fuse_mlp.weight", (int64_t)E * V * E, &Wfuse))) return rc;
    if ((rc = get_vec(h, L + ".fuse_mlp.bias", E, &bfuse))) return rc;
    float *mfold, *M1;
    if ((rc = h->pool.alloc(&mfold, (int64_t)E * FC))) return rc;       // [E][128] = Wv · Wp
    if ((rc = h->pool.alloc(&M1, (int64_t)V * E * E))) return rc;       // per view [E][E] = Wfuse_v · Wop
    if ((rc = small_matmul_ex(Wv, E, Wp, FC, nullptr, mfold, FC, E, FC, E, st))) return rc;
    float* Wc = h->tk_c.f32 + (int64_t)r * E * V * KA;
    const int64_t ldc = (int64_t)V * KA;
    for (int v = 0; v < V; ++v) {
        float* M1v = M1 + (int64_t)v * E * E;
        if ((rc = small_matmul_ex(Wfuse + (int64_t)v * E, (int64_t)V * E, Wop, E, nullptr, M1v, E, E, E, E, st))) return rc;
        float* Wcv = Wc + (int64_t)v * KA;
        for (int hh = 0; hh < TOK_NH; ++hh)      // raw-channel block of head hh: M1[:, head] · mfold[head, :]
            if ((rc = small_matmul_ex(M1v + hh * HD, E, mfold + (int64_t)hh * HD * FC, FC, nullptr, Wcv + hh * TOK_RAWC, ldc, E,
                                      TOK_RAWC, HD, st))) return rc;
        // E-term block (sampled position table P'): M1 itself
        EGR_CUDA_OK(cudaMemcpy2DAsync(Wcv + TOK_NH * TOK_RAWC, sizeof(float) * ldc, M1v, sizeof(float) * E, sizeof(float) * E, E,
                                      cudaMemcpyDeviceToDevice, st));
        // validity column: Wfuse_v · b_op
        if ((rc = small_matmul_ex(Wfuse + (int64_t)v * E, (int64_t)V * E, bop, 1, nullptr, Wcv + TOK_NH * TOK_RAWC + E, ldc, E, 1, E, st)))
            return rc;
    }
    EGR_CUDA_OK(cudaMemcpyAsync(h->tk_c.bias + r * E, bfuse, sizeof(float) * E, cudaMemcpyDeviceToDevice, st));
    return EGR_OK;
}

// Q1 rest + A1 A2 A3 + post_norm over all tokens of G refiner groups (weights from refiner r0 on)
int run_tokens_batched(egr_mvfex* h, int B, int G, int r0, const Bufs& w, const float* bfb, int64_t bfb_bs, int64_t bfb_gs,
                       const float* anchors, const uint8_t* valid, cudaStream_t st) {
    const int J = h->J, V = h->V, E = EMB, KA = h->KA;
    const int T = B * J;
    const float* const* P = h->d_ptrs;
    auto ptr = [&](int which) { return P + which * 4 + r0; };
    int rc;
    // sp: EGR_PREC_FP16 runs the token Linears as "3x TF32" (gemm.cuh `ka`): every GEMM operand row is [x | x_lo] (2K floats),
    // written by its producer kernel, or by tok_make_lo after a GEMM epilogue (`operand` outputs: ldd = 2N)
    const int sp = h->tok3x ? 1 : 0;
    const int tprec = sp ? PREC_TF32X3 : PREC_TF32;
    auto gemm = [&](const float* A, int K, const WMat& W, float* D, int epi, int rnd, bool operand) -> int {
        GemmDesc d;
        const int64_t ldd = (int64_t)W.N * ((sp && operand) ? 2 : 1);
        d.A = A; d.lda = (int64_t)K * (1 + sp); d.M = T; d.D = D; d.ldd = ldd; d.epi = epi; d.round_tf32 = sp ? 0 : rnd;
        d.groups = G; d.a_gs = (int64_t)T * K * (1 + sp); d.d_gs = (int64_t)T * ldd;
        if (int r = run_gemm(d, W, r0, tprec, true, st)) return r;
        return (sp && operand) ? tok_make_lo(D, (int64_t)G * T, W.N, st) : EGR_OK;
    };
    // heatmap_proj.2, + fc_bfb(avgpool) + joint embed, fc_query + ReLU
    if ((rc = gemm(w.q1, E, h->tk_hp2, w.tz, EPI_NONE, 0, false))) return rc;
    if ((rc = tok_avgpool(bfb, bfb_bs, bfb_gs, 64, w.tpool, G, B, 512, st, sp))) return rc;
    {
        GemmDesc d;
        d.A = w.tpool; d.lda = 512 * (1 + sp); d.M = B; d.D = w.tvb; d.ldd = E; d.groups = G; d.a_gs = (int64_t)B * 512 * (1 + sp);
        d.d_gs = (int64_t)B * E;
        if ((rc = run_gemm(d, h->tk_bfb, r0, tprec, true, st))) return rc;
    }
    if ((rc = tok_add_query(w.tz, w.tvb, ptr(TP_JQ), w.to, G, B, J, E, st, sp))) return rc;
    if ((rc = gemm(w.to, E, h->tk_fcq, w.tx, EPI_RELU, 1, true))) return rc;
    // A1: offsets + logits, sampling, folded value/output/fuse GEMM, residual + LN
    if ((rc = gemm(w.tx, E, h->tk_sa, w.toa, EPI_NONE, 0, false))) return rc;
    TokSampleArgs sa{};
    sa.G = G; sa.B = B; sa.V = V; sa.J = J; sa.H = FH; sa.W = FW; sa.E = E; sa.KA = KA; sa.oa = w.toa; sa.anchors = anchors;
    const bool fold16 = sp && h->tk_c.f16x3 != nullptr;      // sampled operand as fp16 pairs [x_hi | x_lo] (PREC_F16X3)
    sa.valid = valid; sa.X = w.Xh; sa.ptab = h->d_ptab16 + r0; sa.A = w.tA; sa.split = fold16 ? 2 : sp;
    if ((rc = tok_sample(sa, act_code(h->prec), st))) return rc;
    if (fold16) {
        GemmDesc d;
        d.A = w.tA; d.lda = (int64_t)V * KA * 2; d.M = T; d.D = w.tz; d.ldd = h->tk_c.N; d.epi = EPI_NONE;
        d.groups = G; d.a_gs = (int64_t)T * V * KA * 2; d.d_gs = (int64_t)T * h->tk_c.N;
        if ((rc = run_gemm(d, h->tk_c, r0, PREC_F16X3, true, st))) return rc;
    } else if ((rc = gemm(w.tA, V * KA, h->tk_c, w.tz, EPI_NONE, 0, false))) return rc;
    if ((rc = tok_add_ln(w.tx, w.tz, w.tx, G, T, E, ptr(TP_LNC_W), ptr(TP_LNC_B), st, sp))) return rc;
    // A2: joint self-attention
    if ((rc = gemm(w.tx, E, h->tk_qkv, w.tqkv, EPI_NONE, 0, false))) return rc;
    if ((rc = tok_attn(w.tqkv, w.to, G * B, J, E, st, sp))) return rc;
    if ((rc = gemm(w.to, E, h->tk_o, w.tz, EPI_NONE, 0, false))) return rc;
    if ((rc = tok_add_ln(w.tx, w.tz, w.tx, G, T, E, ptr(TP_LNS_W), ptr(TP_LNS_B), st, sp))) return rc;
    // A3: FFN
    if ((rc = gemm(w.tx, E, h->tk_f1, w.thid, EPI_GELU, 1, true))) return rc;
    if ((rc = gemm(w.thid, TOK_FF, h->tk_f2, w.tz, EPI_NONE, 0, false))) return rc;
    if ((rc = tok_add_ln(w.tx, w.tz, w.tx, G, T, E, ptr(TP_LNF_W), ptr(TP_LNF_B), st, sp))) return rc;
    // post_norm + token image + T1 (1x1 15->64 ReLU, 1x1 64->128) in one tensor-core kernel
    return tok_head_tc(w.tx, G, B, J, ptr(TP_PN_W), ptr(TP_PN_B), h->t1_0.f32, h->t1_0.bias, h->t1_3.f32, h->t1_3.bias, r0,
                       w.t1, st, E * (1 + sp), h->prec == EGR_PREC_FP16);
}

// the refiner chain for G groups whose weights start at refiner r0
int run_refiners(egr_mvfex* h, int B, int G, int r0, const Bufs& w, const void* Xown, const float* bfb, int64_t bfb_bs,
                 int64_t bfb_gs, const float* anchors, const uint8_t* valid, float* hm_refined, int64_t hm_bs,
                 int64_t hm_gs, float* feat_refined, int64_t ft_bs, int64_t ft_gs, cudaStream_t st) {
    const int prec = h->prec, J = h->J;
    const int bf = (prec == EGR_PREC_BF16);
    const bool f16m = (prec == EGR_PREC_FP16);
    int rc;
    GemmDesc d;
    EGR_MARK("Q1a", st);
    // Q1a: relu(heatmap_proj.0(heatmap))  [G][B*J][4096] -> [G][B*J][256] fp32 (rows [x | x_lo] for the 3x-TF32 token Linears)
    d = GemmDesc();
    const int q1w = EMB * (h->tok3x ? 2 : 1);
    d.A = w.hmT; d.lda = FHW; d.M = B * J; d.D = w.q1; d.ldd = q1w; d.epi = EPI_RELU; d.round_tf32 = (h->tokb && !h->tok3x) ? 1 : 0;
    d.groups = G; d.a_gs = (int64_t)B * J * FHW; d.d_gs = (int64_t)B * J * q1w;
    if ((rc = run_gemm(d, h->hp0, r0, prec, /*out_f32=*/true, st))) return rc;
    if (h->tok3x && (rc = tok_make_lo(w.q1, (int64_t)G * B * J, EMB, st))) return rc;
    EGR_MARK("tokens", st);
    // Q1 rest + A1 A2 A3 + post_norm
    MvfTokenArgs ta{};
    ta.B = B; ta.V = h->V; ta.J = J; ta.H = FH; ta.W = FW; ta.r0 = r0; ta.G = G;
    ta.q1 = w.q1; ta.bfb = bfb; ta.bfb_bs = bfb_bs; ta.bfb_gs = bfb_gs; ta.bfb_hw = 64;
    ta.anchors = anchors; ta.valid = valid; ta.X = w.Xh; ta.xT = w.xT; ta.w = h->d_tokw;
    if (h->tokb) {
        if ((rc = run_tokens_batched(h, B, G, r0, w, bfb, bfb_bs, bfb_gs, anchors, valid, st))) return rc;
    } else {
        if ((rc = launch_mvf_tokens(ta, bf, st))) return rc;
    }
    EGR_MARK("T1", st);
    // T1: 1x1(15->64) ReLU, then the 1x1(64->128) commuted in front of the bilinear x2 (both linear);
    // the batched token path has already produced t1 (tok_head_tc)
    if (!h->tokb) {
    d = GemmDesc();
    d.A = w.xT; d.lda = 16; d.M = B * NPOS; d.D = w.h1t; d.ldd = 64; d.epi = EPI_RELU;
    d.groups = G; d.a_gs = (int64_t)B * NPOS * 16; d.d_gs = (int64_t)B * NPOS * 64;
    {
        GemmDesc t = d; t.N = 64; t.K = 16; t.W = h->t1_0.f32 + (int64_t)r0 * 64 * 16; t.bias = h->t1_0.bias + r0 * 64;
        t.w_gs = 64 * 16; t.b_gs = 64;
        if ((rc = gemm_simt(t, bf, bf, st))) return rc;
        t = GemmDesc();
        t.A = w.h1t; t.lda = 64; t.M = B * NPOS; t.D = w.t1; t.ldd = 128; t.epi = EPI_NONE; t.N = 128; t.K = 64;
        t.groups = G; t.a_gs = (int64_t)B * NPOS * 64; t.d_gs = (int64_t)B * NPOS * 128;
        t.W = h->t1_3.f32 + (int64_t)r0 * 128 * 64; t.bias = h->t1_3.bias + r0 * 128; t.w_gs = 128 * 64; t.b_gs = 128;
        if ((rc = gemm_simt(t, bf, bf, st))) return rc;
    }
    }
    EGR_MARK("F1a", st);
    // F1a: 1x1(128->256) ReLU @64x64
    d = GemmDesc();
    d.A = Xown; d.lda = FC; d.M = B * FHW; d.D = w.a1; d.ldd = 256; d.epi = EPI_RELU;
    d.groups = G; d.a_gs = (int64_t)B * FHW * FC; d.d_gs = (int64_t)B * FHW * 256;
    if ((rc = run_gemm(d, h->f1_0, r0, prec, false, st))) return rc;
    EGR_MARK("F1b", st);
    // F1b: 3x3 s2 (256->512) ReLU -> 32x32
    d = GemmDesc();
    // asplit (EGR_PREC_FP16): the 32x32 activations of the refine path travel as fp16 pairs [hi | lo] (GemmDesc.out_pair) and
    // are consumed as split operands (PREC_F16X3): five of the fourteen 10-bit roundings between the input features and the
    // refined heatmap disappear
    const int ap = h->asplit ? 2 : 1;
    const int cprec = h->asplit ? PREC_F16X3 : prec;      // consumer of a pair tensor
    // the 512-channel F1b output only with asplit >= 2: its pair costs F1b + F1c 0.1 ms (and F1b's epilogue 9 % of its tensor
    // rate) for one rounding out of fourteen
    const int ap1 = h->asplit >= 2 ? 2 : 1;
    d.A = w.a1; d.amode = A_CONV3S2; d.Hin = FH; d.Win = FW; d.Cin = 256; d.M = B * 1024; d.D = w.b1; d.ldd = 512 * ap1;
    d.epi = EPI_RELU; d.groups = G; d.a_gs = (int64_t)B * FHW * 256; d.d_gs = (int64_t)B * 1024 * 512 * ap1; d.out_pair = h->asplit >= 2;
    if ((rc = run_gemm(d, h->f1_2, r0, prec, false, st))) return rc;
    EGR_MARK("F1c", st);
    // F1c: 1x1(512->128) ReLU, fused with "+ offset_pred": + relu(up2(t1))   (:715 offset_pred + frame_feat)
    d = GemmDesc();
    d.A = w.b1; d.lda = 512 * ap1; d.M = B * 1024; d.D = w.ff; d.ldd = 128 * ap; d.epi = EPI_RELU_ADDUP; d.aux = w.t1;
    d.Hout = 32; d.Wout = 32; d.groups = G; d.a_gs = (int64_t)B * 1024 * 512 * ap1; d.d_gs = (int64_t)B * 1024 * 128 * ap;
    d.aux_gs = (int64_t)B * NPOS * 128; d.out_pair = h->asplit ? 1 : 0;
    if ((rc = run_gemm(d, h->f1_4, r0, h->asplit >= 2 ? PREC_F16X3 : prec, false, st))) return rc;
    EGR_MARK("R1a", st);
    // R1a: 1x1(128->128) ReLU @32x32 ; R1b: second 1x1 commuted in front of the upsample
    d = GemmDesc();
    d.A = w.ff; d.lda = 128 * ap; d.M = B * 1024; d.D = w.r1a; d.ldd = 128 * ap; d.epi = EPI_RELU;
    d.groups = G; d.a_gs = d.d_gs = (int64_t)B * 1024 * 128 * ap; d.out_pair = h->asplit ? 1 : 0;
    if ((rc = run_gemm(d, h->r1_0, r0, cprec, false, st))) return rc;
    EGR_MARK("R1b", st);
    // on the tensor-core path the pre-upsample maps z are written in fp16 so that the tails interpolate in half2
    const bool z16 = is16(prec) && g_opt_tc;
    d.A = w.r1a; d.D = w.z; d.ldd = 128; d.d_gs = (int64_t)B * 1024 * 128; d.epi = EPI_NONE; d.out_pair = 0;
    if ((rc = run_gemm(d, h->r1_3, r0, cprec, false, st, z16 ? DT_F16 : DT_BF16))) return rc;
    EGR_MARK("R1tail", st);
    // R1c: relu(up2(.)) -> refined features: fp32 NCHW module output (optional) + channels-last copies for H2 / pose3d
    if ((rc = up2_relu_dual(w.z, z16 ? 2 : bf, B, G, 32, 32, FC, feat_refined, ft_bs, ft_gs, w.refn, f16m ? 2 : bf ? 1 : 3, w.refn_hp,
                            h->export_hp == 2 ? 2 : 0, st, f16m ? 1 : 0))) return rc;
    EGR_MARK("H2a", st);
    // H2a: 3x3 s2 (128->256) ReLU ; H2b: 1x1(256->256) ReLU ; H2c: 1x1(256->128) (commuted) ; tail: up2, ReLU, 1x1->15
    d = GemmDesc();
    d.A = w.refn; d.amode = A_CONV3S2; d.Hin = FH; d.Win = FW; d.Cin = FC; d.M = B * 1024; d.D = w.b1; d.ldd = 256 * ap;
    d.epi = EPI_RELU; d.groups = G; d.a_gs = (int64_t)B * FHW * FC; d.d_gs = (int64_t)B * 1024 * 256 * ap; d.out_pair = h->asplit ? 1 : 0;
    if (!w.refn) {
        // the refined features exist only as the fp16 channels-last copy shared with pose3d: fp16 operands, bf16 output
        d.A = w.refn_hp; d.N = 256; d.K = 9 * FC; d.W = h->h2_0_f16 + (int64_t)r0 * 256 * 9 * FC; d.w_gs = (int64_t)256 * 9 * FC;
        d.bias = h->h2_0.bias + r0 * 256; d.b_gs = 256;
        if ((rc = gemm_tc(d, DT_F16, DT_BF16, st))) return rc;
    } else if ((rc = run_gemm(d, h->h2_0, r0, prec, false, st))) return rc;
    EGR_MARK("H2b", st);
    d = GemmDesc();
    d.A = w.b1; d.lda = 256 * ap; d.M = B * 1024; d.D = w.c1; d.ldd = 256 * ap; d.epi = EPI_RELU;
    d.groups = G; d.a_gs = d.d_gs = (int64_t)B * 1024 * 256 * ap; d.out_pair = h->asplit ? 1 : 0;
    if ((rc = run_gemm(d, h->h2_2, r0, cprec, false, st))) return rc;
    EGR_MARK("H2c", st);
    d = GemmDesc();
    d.A = w.c1; d.lda = 256 * ap; d.M = B * 1024; d.D = w.z; d.ldd = 128; d.epi = EPI_NONE;
    d.groups = G; d.a_gs = (int64_t)B * 1024 * 256 * ap; d.d_gs = (int64_t)B * 1024 * 128;
    if ((rc = run_gemm(d, h->h2_5, r0, cprec, false, st, z16 ? DT_F16 : DT_BF16))) return rc;
    EGR_MARK("H2tail", st);
    int wsel[4] = {r0, r0 + 1, r0 + 2, r0 + 3};
    if ((rc = head_up_conv(w.z, f16m ? 3 : z16 ? 2 : bf, h->h2_7w, h->h2_7b, wsel, B, G, 32, 32, FC, J, hm_refined, hm_bs, hm_gs, nullptr, st)))
        return rc;
    EGR_MARK(nullptr, st);
    return EGR_OK;
}

void note(egr_mvfex* h, const char* name, void* p, int64_t bytes) { h->dbg[name] = std::make_pair(p, bytes); }

void note_all(egr_mvfex* h, const Bufs& w, int B, int G) {
    const int64_t s = is16(h->prec) ? 2 : 4;
    note(h, "Xh", w.Xh, (int64_t)h->V * B * FHW * FC * s);
    note(h, "q1", w.q1, (int64_t)G * B * h->J * EMB * 4 * (h->tok3x ? 2 : 1));
    note(h, "xT", w.xT, (int64_t)G * B * NPOS * 16 * s);
    note(h, "t1", w.t1, (int64_t)G * B * NPOS * 128 * s);
    note(h, "ff", w.ff, (int64_t)G * B * 1024 * 128 * s);
    note(h, "a1", w.a1, (int64_t)G * B * FHW * 256 * s);
    note(h, "hmT", w.hmT, (int64_t)G * B * h->J * FHW * s);
    note(h, "anchors", w.anch, (int64_t)B * h->V * h->J * 8);
    note(h, "valid", w.valid, (int64_t)B * h->V * h->J);
}

}  // namespace

extern "C" int egr_set_option(const char* key, int value) {
    if (key && std::string(key) == "tc") { g_opt_tc = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "pose_p2_fp16") { g_opt_pose_p2_fp16 = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "pose_p2_bf16") { g_opt_pose_p2_bf16 = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "ws") { g_opt_ws = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "pdl") { g_opt_pdl = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "tok_batched") { g_opt_tok_batched = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "wsplit") { g_opt_wsplit = value < 0 ? 0 : value > 2 ? 2 : value; return EGR_OK; }
    if (key && std::string(key) == "tok3x") { g_opt_tok3x = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "stem_fused") { g_opt_stem_fused = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "fpn_epi") { g_opt_fpn_epi = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "conv_prefetch") { g_opt_conv_prefetch = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "pair") { g_opt_pair = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "asplit") { g_opt_asplit = value < 0 ? 0 : value > 2 ? 2 : value; return EGR_OK; }
    if (key && std::string(key) == "fold16") { g_opt_fold16 = value ? 1 : 0; return EGR_OK; }
    if (key && std::string(key) == "tail_mma") { g_opt_tail_mma = value ? 1 : 0; return EGR_OK; }
    return fail(EGR_ERR_INVALID, "unknown option '%s'", key ? key : "(null)");
}

extern "C" int egr_mvfex_create(int num_views, int num_heatmap, float heatmap_threshold, int precision, egr_mvfex** out) {
    EGR_CHECK(out, EGR_ERR_INVALID, "mvfex_create: null out");
    EGR_CHECK(num_views == 2 || num_views == 4, EGR_ERR_UNSUPPORTED,
              "mvfex: num_views=%d (shipped configs use 4, or 2 for stereo-front)", num_views);
    EGR_CHECK(num_heatmap == 15, EGR_ERR_UNSUPPORTED, "mvfex: num_heatmap=%d (shipped configs use 15)", num_heatmap);
    EGR_CHECK(precision == EGR_PREC_FP32 || precision == EGR_PREC_BF16 || precision == EGR_PREC_FP16, EGR_ERR_INVALID,
              "mvfex: precision %d", precision);
    EGR_CHECK(precision != EGR_PREC_FP16 || (g_opt_tc && g_opt_tok_batched), EGR_ERR_UNSUPPORTED,
              "mvfex: EGR_PREC_FP16 is a tensor-core mode (options tc / tok_batched must be on)");
    if (int rc = require_device()) return rc;
    egr_mvfex* h = new egr_mvfex();
    h->V = num_views; h->J = num_heatmap; h->thr = heatmap_threshold; h->prec = precision;
    h->head_sets = (num_views == 2) ? 1 : 2;
    h->tokb = is16(precision) && g_opt_tc && g_opt_tok_batched;
    h->tok3x = precision == EGR_PREC_FP16 && g_opt_tok3x;
    h->asplit = (precision == EGR_PREC_FP16 && g_opt_wsplit >= 1) ? g_opt_asplit : 0;
    h->KA = tok_ka(EMB, true);
    *out = h;
    return EGR_OK;
}

extern "C" int egr_mvfex_destroy(egr_mvfex* h) {
    if (!h) return EGR_OK;
    cudaDeviceSynchronize();
    h->pool.release();
    delete h;
    return EGR_OK;
}

extern "C" int egr_mvfex_set_param(egr_mvfex* h, const char* key, const float* ptr, int64_t numel) {
    EGR_CHECK(h, EGR_ERR_INVALID, "mvfex_set_param: null handle");
    h->packed = false;
    return h->params.set(key, ptr, numel);
}

extern "C" int egr_mvfex_prepack(egr_mvfex* h, void* stream) {
    EGR_CHECK(h, EGR_ERR_INVALID, "mvfex_prepack: null handle");
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    EGR_CUDA_OK(cudaStreamSynchronize(st));
    h->pool.release();
    h->packed = false;
    const int V = h->V, J = h->J;
    int rc;
    if (is16(h->prec) && g_opt_tc) {
        if ((rc = gemm_tc_init())) return rc;
    }
    EGR_CHECK(h->prec != EGR_PREC_FP16 || (g_opt_tc && g_opt_tok_batched), EGR_ERR_UNSUPPORTED,
              "mvfex: EGR_PREC_FP16 is a tensor-core mode (options tc / tok_batched must be on)");
    h->tokb = is16(h->prec) && g_opt_tc && g_opt_tok_batched;
    h->tok3x = h->prec == EGR_PREC_FP16 && g_opt_tok3x;
    h->asplit = (h->prec == EGR_PREC_FP16 && g_opt_wsplit >= 1) ? g_opt_asplit : 0;
    h->KA = tok_ka(EMB, true);
    auto head = [&](const char* sub) { return [sub](int s) { return std::string(kHead[s]) + sub; }; };
    auto ref = [&](const char* sub) { return [sub](int s) { return std::string(kRefiner4[s]) + sub; }; };
    // a standalone HeatmapMVF registers one refiner only; the full module registers heads + all refiners
    h->has_heads = h->params.map.count(std::string(kHead[0]) + ".0.weight") > 0;
    bool any = h->has_heads;
    for (int r = 0; r < 4; ++r) {
        h->has_ref[r] = r < V && h->params.map.count(std::string(kRefiner4[r]) + ".heatmap_proj.0.weight") > 0;
        any = any || h->has_ref[r];
    }
    EGR_CHECK(any, EGR_ERR_STATE, "mvfex_prepack: no parameters registered");
    bool hp[2] = {h->has_heads, h->has_heads};
    const bool* rp = h->has_ref;
    if ((rc = make_wmat(h, h->h1_0, h->head_sets, 128, 128, W_PLAIN, head(".0"), hp, st))) return rc;
    if ((rc = make_wmat(h, h->h1_2, h->head_sets, 256, 9 * 128, W_CONV3, head(".2"), hp, st))) return rc;
    if ((rc = make_wmat(h, h->h1_4, h->head_sets, 256, 256, W_PLAIN, head(".4"), hp, st))) return rc;
    if ((rc = make_wmat(h, h->h1_7, h->head_sets, 128, 256, W_PLAIN, head(".7"), hp, st))) return rc;
    if ((rc = h->pool.alloc(&h->h1_9w, (int64_t)h->head_sets * J * 128))) return rc;
    if ((rc = h->pool.alloc(&h->h1_9b, (int64_t)h->head_sets * J))) return rc;
    for (int s = 0; s < h->head_sets && h->has_heads; ++s) {
        const float* w = h->params.get(std::string(kHead[s]) + ".9.weight", J * 128, &rc); if (!w) return rc;
        const float* b = h->params.get(std::string(kHead[s]) + ".9.bias", J, &rc); if (!b) return rc;
        EGR_CUDA_OK(cudaMemcpyAsync(h->h1_9w + s * J * 128, w, sizeof(float) * J * 128, cudaMemcpyDeviceToDevice, st));
        EGR_CUDA_OK(cudaMemcpyAsync(h->h1_9b + s * J, b, sizeof(float) * J, cudaMemcpyDeviceToDevice, st));
    }
    if ((rc = make_wmat(h, h->hp0, V, EMB, FHW, W_PLAIN, ref(".heatmap_proj.0"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->f1_0, V, 256, 128, W_PLAIN, ref(".frame_feat_proj_layers.0"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->f1_2, V, 512, 9 * 256, W_CONV3, ref(".frame_feat_proj_layers.2"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->f1_4, V, 128, 512, W_PLAIN, ref(".frame_feat_proj_layers.4"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->t1_0, V, 64, 16, W_PAD16, ref(".head_layers.0.head.0"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->t1_3, V, 128, 64, W_PLAIN, ref(".head_layers.0.head.3"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->r1_0, V, 128, 128, W_PLAIN, ref(".frame_feat_refined_proj_layers.0.0"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->r1_3, V, 128, 128, W_PLAIN, ref(".frame_feat_refined_proj_layers.0.3"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->h2_0, V, 256, 9 * 128, W_CONV3, ref(".conv_heatmap_layers.0.0"), rp, st))) return rc;
    if (h->prec == EGR_PREC_BF16 && g_opt_tc) {
        if ((rc = h->pool.alloc(&h->h2_0_f16, (int64_t)V * 256 * 9 * 128))) return rc;
        if ((rc = cast_f16(h->h2_0.f32, h->h2_0_f16, (int64_t)V * 256 * 9 * 128, st))) return rc;
    }
    if ((rc = make_wmat(h, h->h2_2, V, 256, 256, W_PLAIN, ref(".conv_heatmap_layers.0.2"), rp, st))) return rc;
    if ((rc = make_wmat(h, h->h2_5, V, 128, 256, W_PLAIN, ref(".conv_heatmap_layers.0.5"), rp, st))) return rc;
    if (h->asplit) {      // consumers of pair tensors: [W_hi | W_hi | W_lo] against A = [x_hi | x_lo]
        WMat* cons[5] = {&h->f1_4, &h->r1_0, &h->r1_3, &h->h2_2, &h->h2_5};
        for (WMat* m : cons) {
            if ((rc = h->pool.alloc(&m->f16x3, (int64_t)m->sets * m->N * m->K * 3))) return rc;
            if ((rc = split3_f16(m->f32, m->f16x3, (int64_t)m->sets * m->N, m->K, st))) return rc;
        }
    }
    if ((rc = h->pool.alloc(&h->h2_7w, (int64_t)V * J * 128))) return rc;
    if ((rc = h->pool.alloc(&h->h2_7b, (int64_t)V * J))) return rc;
    std::vector<MvfTokenW> tok(V);
    for (int r = 0; r < V; ++r) {
        if (!h->has_ref[r]) continue;
        const std::string p = kRefiner4[r];
        const float* w = h->params.get(p + ".conv_heatmap_layers.0.7.weight", J * 128, &rc); if (!w) return rc;
        const float* b = h->params.get(p + ".conv_heatmap_layers.0.7.bias", J, &rc); if (!b) return rc;
        EGR_CUDA_OK(cudaMemcpyAsync(h->h2_7w + r * J * 128, w, sizeof(float) * J * 128, cudaMemcpyDeviceToDevice, st));
        EGR_CUDA_OK(cudaMemcpyAsync(h->h2_7b + r * J, b, sizeof(float) * J, cudaMemcpyDeviceToDevice, st));
        MvfTokenW& t = tok[r];
        if ((rc = make_T(h, p + ".heatmap_proj.2", EMB, EMB, &t.hp2_T, st))) return rc;
        if ((rc = get_vec(h, p + ".heatmap_proj.2.bias", EMB, &t.hp2_b))) return rc;
        if ((rc = make_T(h, p + ".fc_bfb", EMB, 512, &t.bfb_T, st))) return rc;
        if ((rc = get_vec(h, p + ".fc_bfb.bias", EMB, &t.bfb_b))) return rc;
        if ((rc = get_vec(h, p + ".joint_query_embed.weight", J * EMB, &t.jq))) return rc;
        if ((rc = make_T(h, p + ".fc_query.0", EMB, EMB, &t.fcq_T, st))) return rc;
        if ((rc = get_vec(h, p + ".fc_query.0.bias", EMB, &t.fcq_b))) return rc;
        if ((rc = get_vec(h, p + ".post_norm.0.weight", EMB, &t.pn_w))) return rc;
        if ((rc = get_vec(h, p + ".post_norm.0.bias", EMB, &t.pn_b))) return rc;
        if ((rc = build_layer(h, p + ".transformer_layers.0", p, t.layer, st))) return rc;
    }
    if ((rc = h->pool.alloc(&h->d_tokw, V))) return rc;
    EGR_CUDA_OK(cudaMemcpyAsync(h->d_tokw, tok.data(), sizeof(MvfTokenW) * V, cudaMemcpyHostToDevice, st));
    if (h->tokb) {
        const int E = EMB;
        if ((rc = alloc_wmat(h, h->tk_hp2, V, E, E, st)) || (rc = alloc_wmat(h, h->tk_fcq, V, E, E, st)) ||
            (rc = alloc_wmat(h, h->tk_bfb, V, E, 512, st)) ||
            (rc = alloc_wmat(h, h->tk_sa, V, TOK_OA, E, st)) || (rc = alloc_wmat(h, h->tk_c, V, E, V * h->KA, st)) ||
            (rc = alloc_wmat(h, h->tk_qkv, V, 3 * E, E, st)) || (rc = alloc_wmat(h, h->tk_o, V, E, E, st)) ||
            (rc = alloc_wmat(h, h->tk_f1, V, TOK_FF, E, st)) || (rc = alloc_wmat(h, h->tk_f2, V, E, TOK_FF, st)))
            return rc;
        std::vector<const float*> ptrs(TP_COUNT * 4, nullptr);
        for (int r = 0; r < V; ++r) {
            if (!h->has_ref[r]) continue;
            if ((rc = build_tokb(h, r, kRefiner4[r], st))) return rc;
            const MvfTokenW& t = tok[r];
            ptrs[TP_LNC_W * 4 + r] = t.layer.lnc_w; ptrs[TP_LNC_B * 4 + r] = t.layer.lnc_b;
            ptrs[TP_LNS_W * 4 + r] = t.layer.lns_w; ptrs[TP_LNS_B * 4 + r] = t.layer.lns_b;
            ptrs[TP_LNF_W * 4 + r] = t.layer.lnf_w; ptrs[TP_LNF_B * 4 + r] = t.layer.lnf_b;
            ptrs[TP_PN_W * 4 + r] = t.pn_w; ptrs[TP_PN_B * 4 + r] = t.pn_b;
            ptrs[TP_PTAB * 4 + r] = t.layer.ptab;
            ptrs[TP_BFB_T * 4 + r] = t.bfb_T; ptrs[TP_BFB_B * 4 + r] = t.bfb_b; ptrs[TP_JQ * 4 + r] = t.jq;
        }
        std::vector<const __nv_bfloat16*> pt16(4, nullptr);
        for (int r = 0; r < V; ++r) {
            if (!h->has_ref[r]) continue;
            __nv_bfloat16* t16 = nullptr;      // 16-bit copy in the activation type (fp16 values behind the bf16 pointer type in EGR_PREC_FP16)
            if ((rc = h->pool.alloc(&t16, (int64_t)V * FHW * E))) return rc;
            if (h->prec == EGR_PREC_FP16) { if ((rc = cast_f16(tok[r].layer.ptab, reinterpret_cast<__half*>(t16), (int64_t)V * FHW * E, st))) return rc; }
            else if ((rc = cast_bf16(tok[r].layer.ptab, t16, (int64_t)V * FHW * E, st))) return rc;
            pt16[r] = t16;
        }
        if ((rc = h->pool.alloc(&h->d_ptab16, 4))) return rc;
        EGR_CUDA_OK(cudaMemcpyAsync(h->d_ptab16, pt16.data(), sizeof(void*) * 4, cudaMemcpyHostToDevice, st));
        EGR_CUDA_OK(cudaStreamSynchronize(st));
        WMat* all[9] = {&h->tk_hp2, &h->tk_fcq, &h->tk_bfb, &h->tk_sa, &h->tk_c, &h->tk_qkv, &h->tk_o, &h->tk_f1, &h->tk_f2};
        for (WMat* m : all) {
            if (h->tok3x) {     // [W_hi | W_hi | W_lo] along K, TF32 values in fp32 containers
                if ((rc = h->pool.alloc(&m->f32x3, (int64_t)m->sets * m->N * m->K * 3))) return rc;
                if ((rc = split3_tf32(m->f32, m->f32x3, (int64_t)m->sets * m->N, m->K, st))) return rc;
            } else if ((rc = round_tf32_inplace(m->f32, (int64_t)m->sets * m->N * m->K, st))) return rc;
        }
        if (h->tok3x && g_opt_fold16 && (V * h->KA) % 64 == 0) {
            WMat* m = &h->tk_c;
            if ((rc = h->pool.alloc(&m->f16x3, (int64_t)m->sets * m->N * m->K * 3))) return rc;
            if ((rc = split3_f16(m->f32, m->f16x3, (int64_t)m->sets * m->N, m->K, st))) return rc;
        }
        if ((rc = h->pool.alloc(&h->d_ptrs, TP_COUNT * 4))) return rc;
        EGR_CUDA_OK(cudaMemcpyAsync(h->d_ptrs, ptrs.data(), sizeof(const float*) * TP_COUNT * 4, cudaMemcpyHostToDevice, st));
        EGR_CUDA_OK(cudaStreamSynchronize(st));      // ptrs / tok are host locals
    }
    EGR_CUDA_OK(cudaStreamSynchronize(st));
    h->packed = true;
    return EGR_OK;
}

extern "C" int64_t egr_mvfex_workspace_bytes(egr_mvfex* h, int B) {
    if (!h || B <= 0) return 0;
    return carve(h, B, h->V, true, true, nullptr, 0, nullptr);
}

extern "C" int egr_mvfex_forward(egr_mvfex* h, int B, const float* feat, const float* bfb, const float* heatmap_for_anchor,
                                 float* hm_init, float* hm_refined, float* feat_refined, float* anchors_2d,
                                 uint8_t* anchors_valid, void* workspace, int64_t workspace_bytes, void* stream) {
    EGR_CHECK(h, EGR_ERR_INVALID, "mvfex_forward: null handle");
    if (int rc = require_device()) return rc;
    EGR_CHECK(h->packed, EGR_ERR_STATE, "mvfex_forward: parameters changed or never packed; call egr_mvfex_prepack");
    const void* in_staged = h->in_staged;
    h->in_staged = nullptr;                                   // one-shot hint (egr_mvfex_use_staged_input)
    EGR_CHECK(B > 0 && (feat || in_staged) && bfb && hm_init && hm_refined && workspace, EGR_ERR_INVALID,
              "mvfex_forward: null pointer / empty batch");
    EGR_CHECK(feat_refined || h->export_staged, EGR_ERR_INVALID,
              "mvfex_forward: feat_refined may be NULL only when the channels-last copies are exported (egr_mvfex_export_staged)");
    for (int r = 0; r < h->V; ++r)
        EGR_CHECK(h->has_heads && h->has_ref[r], EGR_ERR_STATE, "mvfex_forward: parameters of the init heads / refiner %d were not registered", r);
    Bufs w;
    const int64_t need = carve(h, B, h->V, true, false, workspace, workspace_bytes, &w);
    EGR_CHECK(need <= workspace_bytes, EGR_ERR_STATE, "mvfex_forward: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    const int V = h->V, J = h->J, prec = h->prec, bf = (prec == EGR_PREC_BF16);
    const bool f16m = (prec == EGR_PREC_FP16);
    int rc;
    gemm_tc_set_scratch(w.splitk);
    EGR_MARK("stage_nhwc", st);
    // S0: NCHW fp32 -> view-major channels-last staging copy; skipped when the producer already wrote that layout
    if (in_staged) w.Xh = const_cast<void*>(in_staged);       // read-only from here on
    else if ((rc = nchw_to_nhwc(feat, w.Xh, B, V, FC, FHW, stage_mode(prec), st))) return rc;
    EGR_MARK("H1a", st);
    // H1: init heads, one group per weight set (front: views 0-1, back: views 2-3), two views per group
    const int G1 = h->head_sets, vpg = V / G1;
    GemmDesc d;
    d.A = w.Xh; d.lda = FC; d.M = vpg * B * FHW; d.D = w.h1a; d.ldd = 128; d.epi = EPI_RELU;
    d.groups = G1; d.a_gs = d.d_gs = (int64_t)vpg * B * FHW * FC;
    if ((rc = run_gemm(d, h->h1_0, 0, prec, false, st))) return rc;
    EGR_MARK("H1b", st);
    d = GemmDesc();
    d.A = w.h1a; d.amode = A_CONV3S2; d.Hin = FH; d.Win = FW; d.Cin = FC; d.M = vpg * B * 1024; d.D = w.b1; d.ldd = 256;
    d.epi = EPI_RELU; d.groups = G1; d.a_gs = (int64_t)vpg * B * FHW * FC; d.d_gs = (int64_t)vpg * B * 1024 * 256;
    if ((rc = run_gemm(d, h->h1_2, 0, prec, false, st))) return rc;
    EGR_MARK("H1c", st);
    d = GemmDesc();
    d.A = w.b1; d.lda = 256; d.M = vpg * B * 1024; d.D = w.c1; d.ldd = 256; d.epi = EPI_RELU;
    d.groups = G1; d.a_gs = d.d_gs = (int64_t)vpg * B * 1024 * 256;
    if ((rc = run_gemm(d, h->h1_4, 0, prec, false, st))) return rc;
    EGR_MARK("H1d", st);
    d = GemmDesc();
    const bool z16 = is16(prec) && g_opt_tc;     // pre-upsample maps in fp16 on the tensor-core path (half2 interpolation in the tails)
    d.A = w.c1; d.lda = 256; d.M = vpg * B * 1024; d.D = w.z; d.ldd = 128; d.epi = EPI_NONE;
    d.groups = G1; d.a_gs = (int64_t)vpg * B * 1024 * 256; d.d_gs = (int64_t)vpg * B * 1024 * 128;
    if ((rc = run_gemm(d, h->h1_7, 0, prec, false, st, z16 ? DT_F16 : DT_BF16))) return rc;
    EGR_MARK("H1tail", st);
    int wsel[4] = {0, 0, 1, 1};
    if (G1 == 1) wsel[2] = wsel[3] = 0;
    if ((rc = head_up_conv(w.z, f16m ? 3 : z16 ? 2 : bf, h->h1_9w, h->h1_9b, wsel, B, V, 32, 32, FC, J, hm_init, (int64_t)V * J * FHW,
                           (int64_t)J * FHW, w.hmT, st))) return rc;
    EGR_MARK("D1", st);
    // D1: anchors from heatmap_for_anchor when given (:293-296), else from the init heatmap
    const float* src = heatmap_for_anchor ? heatmap_for_anchor : hm_init;
    float* anch = anchors_2d ? anchors_2d : w.anch;
    uint8_t* val = anchors_valid ? anchors_valid : w.valid;
    if ((rc = egr_decode_argmax(src, (int64_t)B * V, J, FH, FW, h->thr, 1, anch, w.maxv, val, nullptr, stream))) return rc;
    // refiners (one group per view)
    rc = run_refiners(h, B, V, 0, w, w.Xh, bfb, (int64_t)V * 512 * 64, (int64_t)512 * 64, anch, val, hm_refined,
                      (int64_t)V * J * FHW, (int64_t)J * FHW, feat_refined, (int64_t)V * FC * FHW, (int64_t)FC * FHW, st);
    if (rc) return rc;
    note_all(h, w, B, V);
    h->st_init = w.Xh; h->st_refined = w.refn; h->st_refined_hp = f16m ? w.refn : w.refn_hp;
    return EGR_OK;
}

extern "C" int egr_mvfex_use_staged_input(egr_mvfex* h, const void* feat_vmajor_nhwc_bf16) {
    EGR_CHECK(h, EGR_ERR_INVALID, "mvfex_use_staged_input: null handle");
    EGR_CHECK(is16(h->prec), EGR_ERR_UNSUPPORTED, "mvfex_use_staged_input: needs the bf16 precision (the staged copy is bf16) or the fp16 precision (fp16 copy)");
    EGR_CHECK(((uintptr_t)feat_vmajor_nhwc_bf16 & 15) == 0, EGR_ERR_INVALID, "mvfex_use_staged_input: pointer must be 16-byte aligned");
    h->in_staged = feat_vmajor_nhwc_bf16;
    return EGR_OK;
}

extern "C" int egr_mvfex_export_staged(egr_mvfex* h, int enable) {
    EGR_CHECK(h, EGR_ERR_INVALID, "mvfex_export_staged: null handle");
    EGR_CHECK(enable >= 0 && enable <= 4, EGR_ERR_INVALID, "mvfex_export_staged: mode %d", enable);
    EGR_CHECK(enable != 4 || (is16(h->prec) && g_opt_tc), EGR_ERR_UNSUPPORTED, "mvfex_export_staged: mode 4 needs the tensor-core path");
    h->export_staged = enable != 0;
    h->export_hp = (enable == 1) ? 1 : (enable >= 3) ? 2 : 0;
    h->refn_f16_only = enable == 4;
    h->st_init = h->st_refined = nullptr; h->st_refined_hp = nullptr;
    return EGR_OK;
}

extern "C" int egr_mvfex_staged(egr_mvfex* h, const void** init_nhwc, const void** refined_nhwc, const void** refined_nhwc_hp,
                                int* act_is_bf16) {
    EGR_CHECK(h && init_nhwc && refined_nhwc && refined_nhwc_hp && act_is_bf16, EGR_ERR_INVALID, "mvfex_staged: null argument");
    EGR_CHECK(h->st_init, EGR_ERR_STATE, "mvfex_staged: no forward has run since export was enabled");
    *init_nhwc = h->st_init; *refined_nhwc = h->st_refined; *refined_nhwc_hp = h->st_refined_hp;
    *act_is_bf16 = act_code(h->prec);       // 0 fp32, 1 bf16, 2 fp16
    return EGR_OK;
}

extern "C" int egr_mvfex_refiner_forward(egr_mvfex* h, int r, int B, const float* heatmap, const float* frame_feat,
                                         const float* feat_mv, const float* anchors_2d, const uint8_t* anchors_valid,
                                         const float* bfb, float* hm_refined, float* feat_refined, void* workspace,
                                         int64_t workspace_bytes, void* stream) {
    EGR_CHECK(h, EGR_ERR_INVALID, "mvfex_refiner_forward: null handle");
    if (int rc = require_device()) return rc;
    EGR_CHECK(h->packed, EGR_ERR_STATE, "mvfex_refiner_forward: call egr_mvfex_prepack first");
    EGR_CHECK(r >= 0 && r < h->V, EGR_ERR_INVALID, "mvfex_refiner_forward: refiner %d of %d", r, h->V);
    EGR_CHECK(h->has_ref[r], EGR_ERR_STATE, "mvfex_refiner_forward: parameters of refiner %d were not registered", r);
    EGR_CHECK(B > 0 && heatmap && frame_feat && feat_mv && anchors_2d && anchors_valid && bfb && hm_refined &&
              feat_refined && workspace, EGR_ERR_INVALID, "mvfex_refiner_forward: null pointer / empty batch");
    Bufs w;
    const int64_t need = carve(h, B, 1, false, true, workspace, workspace_bytes, &w);
    EGR_CHECK(need <= workspace_bytes, EGR_ERR_STATE, "mvfex_refiner_forward: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    const int J = h->J;
    int rc;
    gemm_tc_set_scratch(w.splitk);
    if ((rc = nchw_to_nhwc(feat_mv, w.Xh, B, h->V, FC, FHW, stage_mode(h->prec), st))) return rc;
    if ((rc = nchw_to_nhwc(frame_feat, w.Xown, B, 1, FC, FHW, stage_mode(h->prec), st))) return rc;
    if ((rc = cast_act(heatmap, w.hmT, act_code(h->prec), (int64_t)B * J * FHW, st))) return rc;
    rc = run_refiners(h, B, 1, r, w, w.Xown, bfb, (int64_t)512 * 64, 0, anchors_2d, anchors_valid, hm_refined,
                      (int64_t)J * FHW, 0, feat_refined, (int64_t)FC * FHW, 0, st);
    if (rc) return rc;
    note_all(h, w, B, 1);
    return EGR_OK;
}

extern "C" int egr_mvfex_debug_buffer(egr_mvfex* h, const char* name, void** ptr, int64_t* bytes) {
    EGR_CHECK(h && name && ptr && bytes, EGR_ERR_INVALID, "debug_buffer: null argument");
    auto it = h->dbg.find(name);
    EGR_CHECK(it != h->dbg.end(), EGR_ERR_INVALID, "debug_buffer: unknown buffer '%s'", name);
    *ptr = it->second.first;
    *bytes = it->second.second;
    return EGR_OK;
}
