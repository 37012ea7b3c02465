// pose3d engine: EgoPoseFormerPose3D.forward (estimator/egoposeformer_mvf_ex.py:422-452), conv-MLP proposal branch.
//   P2  conv_frame_feat (:229-239) + mlp_pred (:241-253) on frame_feats_final  -> proposal [B,16,3]
//   P3  fisheye reprojection (in the token kernel)   P1+P4  feat_proj folded into the deformable sampling,
//   3 transformer layers + post_norm + reg_mlp       (token kernel)
// Activations are channels-last, view-major ([V][B][H*W][C]); the flatten "(b v) c h w -> b (v c h w)" of :317
// becomes a K-split GEMM over the 4 per-view [B][64*128] blocks with the weight columns permuted once at prepack.
#include "engine_common.cuh"
#include "token_kernels.cuh"
#include "token_batched.cuh"

namespace egr {
extern int g_opt_tok_batched;
CamCalib make_calib(int cam_id, const float* calib_host);
// P2 (conv_frame_feat + mlp_pred) feeds the 3D proposal directly: bf16 operands cost ~0.1 mm MPJPE on their own
// (measured), which is the whole parity budget, so in EGR_PREC_BF16 the branch keeps fp32 activations and weights and
// multiplies them as TF32 on the tensor cores (PREC_TF32); option "pose_p2_bf16" forces bf16 operands instead.
int g_opt_pose_p2_bf16 = 0;
int g_opt_pose_p2_fp16 = 1;    // EGR_PREC_BF16: fp16 instead of fp32/TF32 activations and weights in P2 (same 10-bit mantissa)
}
using namespace egr;

constexpr int PH = 64, PW = 64, PHW = 4096, PC = 128, PE = 128;

struct PoseTokB {       // token GEMM weights of one layer (fp32 rounded to TF32)
    WMat sa, c, qkv, o, f1, f2, r0;
};

struct egr_pose3d {
    int V = 4, J = 16, L = 3, cam_model = 0, use_init = 1, prec = EGR_PREC_FP32;
    bool packed = false;
    ParamTable params;
    DevPool pool;
    CamCalib cam[4];
    int cam_id[4] = {0, 1, 2, 3};
    WMat c0, c2, c5, c7;          // conv_frame_feat.{0,2,5,7}
    WMat m0, m1, m2;              // mlp_pred.0.0 (permuted), mlp_pred.1.0, mlp_pred.2
    PoseTokenW* d_w = nullptr;
    // one-shot staged inputs for the next forward (egr_pose3d_use_staged)
    const void* st_sampled = nullptr;
    int st_sampled_bf16 = 0;
    const float* st_final_tf32 = nullptr;
    const void* st_final_f16 = nullptr;
    // batched token path (bf16 precision)
    bool tokb = false;
    int KA = 0;
    PoseTokenW tw_host;
    WMat tk_g2, tk_g4;
    PoseTokB tb[4];
    const float** d_ptrs = nullptr;     // device [L][8]: lnc w,b  lns w,b  lnf w,b  post_norm w,b
    std::unordered_map<std::string, std::pair<void*, int64_t>> dbg;
};

namespace {

inline int p2_prec(const egr_pose3d* h) {
    if (!is16(h->prec)) return EGR_PREC_FP32;
    if (h->prec == EGR_PREC_BF16 && g_opt_pose_p2_bf16) return EGR_PREC_BF16;
    if (!g_opt_tc) return EGR_PREC_FP32;
    if (h->prec == EGR_PREC_FP16) return PREC_FP16;
    return g_opt_pose_p2_fp16 ? PREC_FP16 : PREC_TF32;
}

int p_make_wmat(egr_pose3d* h, WMat& m, int N, int K, int kind /*0 plain 1 conv3 2 mlp-permute*/, const std::string& key,
                cudaStream_t st, bool tf32_operand = true, bool f16_operand = true) {
    m.N = N; m.K = K; m.sets = 1;
    int rc = EGR_OK;
    if ((rc = h->pool.alloc(&m.f32, (int64_t)N * K))) return rc;
    if ((rc = h->pool.alloc(&m.bias, N))) return rc;
    const float* w = h->params.get(key + ".weight", (int64_t)N * K, &rc);
    if (!w) return rc;
    if (kind == 0) EGR_CUDA_OK(cudaMemcpyAsync(m.f32, w, sizeof(float) * N * K, cudaMemcpyDeviceToDevice, st));
    else if (kind == 1) { if ((rc = repack_conv3(w, m.f32, N, K / 9, st))) return rc; }
    else { if ((rc = permute_mlp_weight(w, m.f32, N, h->V, PC, 64, st))) return rc; }
    const float* b = h->params.get(key + ".bias", N, &rc);
    if (!b) return rc;
    EGR_CUDA_OK(cudaMemcpyAsync(m.bias, b, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
    if (p2_prec(h) == EGR_PREC_BF16) {
        if ((rc = h->pool.alloc(&m.bf16, (int64_t)N * K))) return rc;
        if ((rc = cast_bf16(m.f32, m.bf16, (int64_t)N * K, st))) return rc;
    } else if (p2_prec(h) == PREC_FP16 && f16_operand) {
        if ((rc = h->pool.alloc(&m.f16, (int64_t)N * K))) return rc;
        if ((rc = cast_f16(m.f32, m.f16, (int64_t)N * K, st))) return rc;
    } else if (p2_prec(h) != EGR_PREC_FP32 && tf32_operand) {      // TF32 stage (also mlp_pred.1 of the fp16 branch)
        if ((rc = round_tf32_inplace(m.f32, (int64_t)N * K, st))) return rc;
    }
    return EGR_OK;
}

int p_make_T(egr_pose3d* h, const std::string& key, int N, int K, const float** out, cudaStream_t st) {
    int rc = EGR_OK;
    const float* w = h->params.get(key + ".weight", (int64_t)N * K, &rc);
    if (!w) return rc;
    float* t = nullptr;
    if ((rc = h->pool.alloc(&t, (int64_t)N * K))) return rc;
    if ((rc = transpose2d(w, t, N, K, st))) return rc;
    *out = t;
    return EGR_OK;
}
int p_vec(egr_pose3d* h, const std::string& key, int n, const float** out) {
    int rc = EGR_OK;
    *out = h->params.get(key, n, &rc);
    return rc;
}

int p_build_layer(egr_pose3d* h, const std::string& p, LayerW& L, cudaStream_t st) {
    const int E = PE, V = h->V;
    int rc;
#define T_(field, name, N, K) if ((rc = p_make_T(h, p + name, N, K, &L.field, st))) return rc
#define V_(field, name, n) if ((rc = p_vec(h, p + name, n, &L.field))) return rc
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
    // fold feat_proj (P1, :144/:431) into the sampling: mfold_T [128][E] = Wf^T · Wv^T, bfold = Wv·bf + bv
    const float *Wv, *bv, *Wf, *bfp;
    if ((rc = p_vec(h, p + ".cross_attn.value_proj.weight", E * E, &Wv))) return rc;
    if ((rc = p_vec(h, p + ".cross_attn.value_proj.bias", E, &bv))) return rc;
    if ((rc = p_vec(h, "feat_proj.weight", E * PC, &Wf))) return rc;
    if ((rc = p_vec(h, "feat_proj.bias", E, &bfp))) return rc;
    float *WvT, *WfT, *mf, *bfold;
    if ((rc = h->pool.alloc(&WvT, E * E))) return rc;
    if ((rc = h->pool.alloc(&WfT, PC * E))) return rc;
    if ((rc = h->pool.alloc(&mf, PC * E))) return rc;
    if ((rc = h->pool.alloc(&bfold, E))) return rc;
    if ((rc = transpose2d(Wv, WvT, E, E, st))) return rc;
    if ((rc = transpose2d(Wf, WfT, E, PC, st))) return rc;
    if ((rc = small_matmul(WfT, WvT, nullptr, mf, PC, E, E, st))) return rc;
    if ((rc = small_matmul(bfp, WvT, bv, bfold, 1, E, E, st))) return rc;
    L.mfold_T = mf;
    L.ptab = nullptr;
    L.bfold = bfold;
    return EGR_OK;
}

struct PBufs {
    void *Xi, *Xf, *p0, *p2, *p3, *p5, *p7;
    float *m0, *m1, *anch;
    uint8_t* valid;
    float *tx, *tz, *toa, *tA, *tqkv, *to, *thid, *tp3;     // batched token path
    float* splitk;                                          // split-K partial sums
};

int64_t p_carve(const egr_pose3d* h, int B, void* base, int64_t cap, PBufs* o) {
    const int64_t si = is16(h->prec) ? 2 : 4;          // sampled map
    const int64_t s = (p2_prec(h) == EGR_PREC_BF16 || p2_prec(h) == PREC_FP16) ? 2 : 4;        // proposal branch
    const int64_t VB = (int64_t)h->V * B;
    Carver c(base, cap);
    PBufs b{};
    b.Xi = c.take(VB * PHW * PC * si);
    b.Xf = c.take(VB * PHW * PC * s);
    b.p0 = c.take(VB * PHW * 64 * s);       // 1x1 128->64 @64x64
    b.p2 = c.take(VB * 1024 * 128 * s);     // 3x3 s2 64->128 @32x32
    b.p3 = c.take(VB * 256 * 128 * s);      // maxpool @16x16
    b.p5 = c.take(VB * 256 * 64 * s);       // 1x1 128->64 @16x16
    b.p7 = c.take(VB * 64 * 128 * s);       // 3x3 s2 64->128 @8x8
    b.m0 = (float*)c.take((int64_t)B * 2048 * 4);
    b.m1 = (float*)c.take((int64_t)B * 128 * 4);
    b.anch = (float*)c.take(VB * h->J * 2 * 4);
    b.valid = (uint8_t*)c.take(VB * h->J);
    if (h->tokb) {
        const int64_t T = (int64_t)B * h->J;
        b.tx = (float*)c.take(T * PE * 4);
        b.tz = (float*)c.take(T * PE * 4);
        b.toa = (float*)c.take(T * TOK_OA * 4);
        b.tA = (float*)c.take(T * h->V * h->KA * 4);
        b.tqkv = (float*)c.take(T * 3 * PE * 4);
        b.to = (float*)c.take(T * PE * 4);
        b.thid = (float*)c.take(T * TOK_FF * 4);
        b.tp3 = (float*)c.take(T * 4 * 4);
    }
    b.splitk = (float*)c.take(SPLITK_SCRATCH_BYTES);
    if (o) *o = b;
    return c.off + 256;
}

// ---- batched token path (EGR_PREC_BF16): derived weights + orchestration (algebra in token_batched.cuh) ----
int p_alloc_wmat(egr_pose3d* h, WMat& m, int N, int K, cudaStream_t st) {
    m.N = N; m.K = K; m.sets = 1; m.bf16 = nullptr;
    if (int rc = h->pool.alloc(&m.f32, (int64_t)N * K)) return rc;
    if (int rc = h->pool.alloc(&m.bias, N)) return rc;
    EGR_CUDA_OK(cudaMemsetAsync(m.f32, 0, sizeof(float) * N * K, st));
    EGR_CUDA_OK(cudaMemsetAsync(m.bias, 0, sizeof(float) * N, st));
    return EGR_OK;
}
int p_copy_rows(egr_pose3d* h, const std::string& key, float* dst_w, float* dst_b, int N, int K, cudaStream_t st) {
    int rc = EGR_OK;
    const float* w = h->params.get(key + ".weight", (int64_t)N * K, &rc);
    if (!w) return rc;
    const float* b = h->params.get(key + ".bias", N, &rc);
    if (!b) return rc;
    EGR_CUDA_OK(cudaMemcpyAsync(dst_w, w, sizeof(float) * N * K, cudaMemcpyDeviceToDevice, st));
    EGR_CUDA_OK(cudaMemcpyAsync(dst_b, b, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
    return EGR_OK;
}

int p_build_tokb_layer(egr_pose3d* h, int l, const LayerW& lw, cudaStream_t st) {
    const int E = PE, V = h->V, KA = h->KA, HD = E / TOK_NH;
    const std::string L = "layers." + std::to_string(l);
    PoseTokB& t = h->tb[l];
    int rc;
    if ((rc = p_alloc_wmat(h, t.sa, TOK_OA, E, st)) || (rc = p_alloc_wmat(h, t.c, E, V * KA, st)) ||
        (rc = p_alloc_wmat(h, t.qkv, 3 * E, E, st)) || (rc = p_alloc_wmat(h, t.o, E, E, st)) ||
        (rc = p_alloc_wmat(h, t.f1, TOK_FF, E, st)) || (rc = p_alloc_wmat(h, t.f2, E, TOK_FF, st)) ||
        (rc = p_alloc_wmat(h, t.r0, E, E, st)))
        return rc;
    if ((rc = p_copy_rows(h, L + ".cross_attn.sampling_offsets", t.sa.f32, t.sa.bias, TOK_NH * TOK_P * 2, E, st))) return rc;
    if ((rc = p_copy_rows(h, L + ".cross_attn.attention_weights", t.sa.f32 + (int64_t)TOK_NH * TOK_P * 2 * E,
                          t.sa.bias + TOK_NH * TOK_P * 2, TOK_NH * TOK_P, E, st))) return rc;
    if ((rc = p_copy_rows(h, L + ".spatial_attn.q_proj", t.qkv.f32, t.qkv.bias, E, E, st))) return rc;
    if ((rc = p_copy_rows(h, L + ".spatial_attn.k_proj", t.qkv.f32 + (int64_t)E * E, t.qkv.bias + E, E, E, st))) return rc;
    if ((rc = p_copy_rows(h, L + ".spatial_attn.v_proj", t.qkv.f32 + (int64_t)2 * E * E, t.qkv.bias + 2 * E, E, E, st))) return rc;
    if ((rc = p_copy_rows(h, L + ".spatial_attn.out_proj", t.o.f32, t.o.bias, E, E, st))) return rc;
    if ((rc = p_copy_rows(h, L + ".ffn.layers.0.0", t.f1.f32, t.f1.bias, TOK_FF, E, st))) return rc;
    if ((rc = p_copy_rows(h, L + ".ffn.layers.1", t.f2.f32, t.f2.bias, E, TOK_FF, st))) return rc;
    if ((rc = p_copy_rows(h, "reg_mlp." + std::to_string(l) + ".0", t.r0.f32, t.r0.bias, E, E, st))) return rc;
    const float *Wv, *Wf, *Wop, *bop, *Wfuse, *bfuse;
    if ((rc = p_vec(h, L + ".cross_attn.value_proj.weight", E * E, &Wv))) return rc;
    if ((rc = p_vec(h, "feat_proj.weight", E * PC, &Wf))) return rc;
    if ((rc = p_vec(h, L + ".cross_attn.output_proj.weight", E * E, &Wop))) return rc;
    if ((rc = p_vec(h, L + ".cross_attn.output_proj.bias", E, &bop))) return rc;
    if ((rc = p_vec(h, L + ".fuse_mlp.weight", (int64_t)E * V * E, &Wfuse))) return rc;
    if ((rc = p_vec(h, L + ".fuse_mlp.bias", E, &bfuse))) return rc;
    float *mfold, *M1;
    if ((rc = h->pool.alloc(&mfold, (int64_t)E * PC))) return rc;       // [E][128] = Wv · Wf
    if ((rc = h->pool.alloc(&M1, (int64_t)V * E * E))) return rc;
    if ((rc = small_matmul_ex(Wv, E, Wf, PC, nullptr, mfold, PC, E, PC, E, st))) return rc;
    const int64_t ldc = (int64_t)V * KA;
    for (int v = 0; v < V; ++v) {
        float* M1v = M1 + (int64_t)v * E * E;
        if ((rc = small_matmul_ex(Wfuse + (int64_t)v * E, (int64_t)V * E, Wop, E, nullptr, M1v, E, E, E, E, st))) return rc;
        float* Wcv = t.c.f32 + (int64_t)v * KA;
        for (int hh = 0; hh < TOK_NH; ++hh) {
            if ((rc = small_matmul_ex(M1v + hh * HD, E, mfold + (int64_t)hh * HD * PC, PC, nullptr, Wcv + hh * TOK_RAWC, ldc, E,
                                      TOK_RAWC, HD, st))) return rc;
            // in-map bilinear weight mass of head hh times (Wv·bf + bv) restricted to the head's channels
            if ((rc = small_matmul_ex(M1v + hh * HD, E, lw.bfold + hh * HD, 1, nullptr, Wcv + TOK_NH * TOK_RAWC + hh, ldc, E, 1, HD, st)))
                return rc;
        }
        if ((rc = small_matmul_ex(Wfuse + (int64_t)v * E, (int64_t)V * E, bop, 1, nullptr, Wcv + TOK_NH * TOK_RAWC + TOK_NH, ldc, E, 1, E, st)))
            return rc;
    }
    EGR_CUDA_OK(cudaMemcpyAsync(t.c.bias, bfuse, sizeof(float) * E, cudaMemcpyDeviceToDevice, st));
    WMat* all[7] = {&t.sa, &t.c, &t.qkv, &t.o, &t.f1, &t.f2, &t.r0};
    for (WMat* m : all)
        if ((rc = round_tf32_inplace(m->f32, (int64_t)m->N * m->K, st))) return rc;
    return EGR_OK;
}

int p_run_tokens_batched(egr_pose3d* h, int B, const PBufs& w, const void* Xs, int bfs, const float* ctm, float* preds,
                         const PoseTokenW& tw, cudaStream_t st) {
    const int J = h->J, V = h->V, E = PE, KA = h->KA, T = B * J;
    int rc;
    auto gemm = [&](const float* A, int K, const WMat& W, float* D, int epi, int rnd) {
        GemmDesc d;
        d.A = A; d.lda = K; d.M = T; d.D = D; d.ldd = W.N; d.epi = epi; d.round_tf32 = rnd;
        return run_gemm(d, W, 0, PREC_TF32, true, st);
    };
    PoseQueryArgs qa{};
    qa.B = B; qa.V = V; qa.J = J; qa.E = E; qa.is_rw = (h->cam_model & 1);
    for (int v = 0; v < 4; ++v) { qa.cam_id[v] = h->cam_id[v]; qa.cam[v] = h->cam[v]; }
    qa.ctm = ctm; qa.mlp_pred = preds; qa.g0_T = tw.g0_T; qa.g0_b = tw.g0_b;
    qa.anchors = w.anch; qa.valid = w.valid; qa.p3 = w.tp3; qa.x0 = w.to;
    if ((rc = pose_query0(qa, st))) return rc;
    if ((rc = gemm(w.to, E, h->tk_g2, w.tz, EPI_RELU, 1))) return rc;
    if ((rc = gemm(w.tz, E, h->tk_g4, w.tx, EPI_NONE, 1))) return rc;
    for (int l = 0; l < h->L; ++l) {
        const PoseTokB& t = h->tb[l];
        const float* const* P = h->d_ptrs + l * 8;
        if ((rc = gemm(w.tx, E, t.sa, w.toa, EPI_NONE, 0))) return rc;
        TokSampleArgs sa{};
        sa.G = 1; sa.B = B; sa.V = V; sa.J = J; sa.H = PH; sa.W = PW; sa.E = E; sa.KA = KA; sa.oa = w.toa; sa.anchors = w.anch;
        sa.valid = w.valid; sa.X = Xs; sa.ptab = nullptr; sa.A = w.tA;
        if ((rc = tok_sample(sa, bfs, st))) return rc;
        if ((rc = gemm(w.tA, V * KA, t.c, w.tz, EPI_NONE, 0))) return rc;
        if ((rc = tok_add_ln(w.tx, w.tz, w.tx, 1, T, E, P + 0, P + 1, st))) return rc;
        if ((rc = gemm(w.tx, E, t.qkv, w.tqkv, EPI_NONE, 0))) return rc;
        if ((rc = tok_attn(w.tqkv, w.to, B, J, E, st))) return rc;
        if ((rc = gemm(w.to, E, t.o, w.tz, EPI_NONE, 0))) return rc;
        if ((rc = tok_add_ln(w.tx, w.tz, w.tx, 1, T, E, P + 2, P + 3, st))) return rc;
        if ((rc = gemm(w.tx, E, t.f1, w.thid, EPI_GELU, 1))) return rc;
        if ((rc = gemm(w.thid, TOK_FF, t.f2, w.tz, EPI_NONE, 0))) return rc;
        // FFN residual + norm, and post_norm[l] of the result in the same launch -> reg_mlp[l] -> + anchors (after the in-place quirk)
        if ((rc = tok_add_ln(w.tx, w.tz, w.tx, 1, T, E, P + 4, P + 5, st, 0, w.to, P + 6, P + 7))) return rc;
        if ((rc = gemm(w.to, E, t.r0, w.tz, EPI_GELU, 0))) return rc;
        if ((rc = pose_reg_out(w.tz, tw.r2_T[l], tw.r2_b[l], w.tp3, preds + (int64_t)(l + 1) * T * 3, T, E, st))) return rc;
    }
    return EGR_OK;
}

}  // namespace

extern "C" int egr_pose3d_create(int num_views, int num_joints, int num_layers, int camera_model,
                                 int use_pred_heatmap_init, int precision, const float* calib_host, egr_pose3d** out) {
    EGR_CHECK(out, EGR_ERR_INVALID, "pose3d_create: null out");
    EGR_CHECK(num_joints == 16, EGR_ERR_UNSUPPORTED, "pose3d: num_joints=%d (shipped configs use 16)", num_joints);
    EGR_CHECK(num_layers >= 1 && num_layers <= 4, EGR_ERR_UNSUPPORTED, "pose3d: num_former_layers=%d", num_layers);
    EGR_CHECK(camera_model >= 0 && camera_model <= 5, EGR_ERR_INVALID, "Unknown camera model !");
    const int need_v = (camera_model <= 1) ? 4 : 2;
    EGR_CHECK(num_views == need_v, EGR_ERR_INVALID, "pose3d: camera model %d needs num_views == %d", camera_model, need_v);
    EGR_CHECK(precision == EGR_PREC_FP32 || precision == EGR_PREC_BF16 || precision == EGR_PREC_FP16, EGR_ERR_INVALID,
              "pose3d: precision %d", precision);
    EGR_CHECK(precision != EGR_PREC_FP16 || (g_opt_tc && g_opt_tok_batched), EGR_ERR_UNSUPPORTED,
              "pose3d: EGR_PREC_FP16 is a tensor-core mode (options tc / tok_batched must be on)");
    if (int rc = require_device()) return rc;
    egr_pose3d* h = new egr_pose3d();
    h->V = num_views; h->J = num_joints; h->L = num_layers; h->cam_model = camera_model;
    h->use_init = use_pred_heatmap_init; h->prec = precision;
    h->tokb = is16(precision) && g_opt_tc && g_opt_tok_batched;
    h->KA = tok_ka(PE, false);
    const int first = (camera_model >= 4) ? 2 : 0;   // stereo_back rigs start at back_left
    for (int v = 0; v < 4; ++v) {
        h->cam_id[v] = (v < num_views) ? first + v : 0;
        h->cam[v] = make_calib(h->cam_id[v], calib_host);
    }
    *out = h;
    return EGR_OK;
}

extern "C" int egr_pose3d_destroy(egr_pose3d* h) {
    if (!h) return EGR_OK;
    cudaDeviceSynchronize();
    h->pool.release();
    delete h;
    return EGR_OK;
}

extern "C" int egr_pose3d_set_param(egr_pose3d* h, const char* key, const float* ptr, int64_t numel) {
    EGR_CHECK(h, EGR_ERR_INVALID, "pose3d_set_param: null handle");
    h->packed = false;
    return h->params.set(key, ptr, numel);
}

extern "C" int egr_pose3d_prepack(egr_pose3d* h, void* stream) {
    EGR_CHECK(h, EGR_ERR_INVALID, "pose3d_prepack: null handle");
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    EGR_CUDA_OK(cudaStreamSynchronize(st));
    h->pool.release();
    h->packed = false;
    int rc;
    if (p2_prec(h) != EGR_PREC_FP32 && g_opt_tc) {
        if ((rc = gemm_tc_init())) return rc;
    }
    if ((rc = p_make_wmat(h, h->c0, 64, 128, 0, "conv_frame_feat.0", st))) return rc;
    if ((rc = p_make_wmat(h, h->c2, 128, 9 * 64, 1, "conv_frame_feat.2", st))) return rc;
    if ((rc = p_make_wmat(h, h->c5, 64, 128, 0, "conv_frame_feat.5", st))) return rc;
    if ((rc = p_make_wmat(h, h->c7, 128, 9 * 64, 1, "conv_frame_feat.7", st))) return rc;
    const int K0 = h->V * PC * 64;
    if ((rc = p_make_wmat(h, h->m0, K0 / 16, K0, 2, "mlp_pred.0.0", st))) return rc;
    if ((rc = p_make_wmat(h, h->m1, K0 / 256, K0 / 16, 0, "mlp_pred.1.0", st, true, false))) return rc;      // TF32 in both branches
    if ((rc = p_make_wmat(h, h->m2, 3 * h->J, K0 / 256, 0, "mlp_pred.2", st, false, false))) return rc;     // fp32 SIMT
    PoseTokenW tw{};
    if ((rc = p_make_T(h, "query_gen_mlp.0", PE, 4, &tw.g0_T, st))) return rc;
    if ((rc = p_vec(h, "query_gen_mlp.0.bias", PE, &tw.g0_b))) return rc;
    if ((rc = p_make_T(h, "query_gen_mlp.2", PE, PE, &tw.g2_T, st))) return rc;
    if ((rc = p_vec(h, "query_gen_mlp.2.bias", PE, &tw.g2_b))) return rc;
    if ((rc = p_make_T(h, "query_gen_mlp.4", PE, PE, &tw.g4_T, st))) return rc;
    if ((rc = p_vec(h, "query_gen_mlp.4.bias", PE, &tw.g4_b))) return rc;
    for (int l = 0; l < h->L; ++l) {
        const std::string li = std::to_string(l);
        if ((rc = p_build_layer(h, "layers." + li, tw.layer[l], st))) return rc;
        if ((rc = p_vec(h, "post_norm." + li + ".weight", PE, &tw.pn_w[l]))) return rc;
        if ((rc = p_vec(h, "post_norm." + li + ".bias", PE, &tw.pn_b[l]))) return rc;
        if ((rc = p_make_T(h, "reg_mlp." + li + ".0", PE, PE, &tw.r0_T[l], st))) return rc;
        if ((rc = p_vec(h, "reg_mlp." + li + ".0.bias", PE, &tw.r0_b[l]))) return rc;
        if ((rc = p_make_T(h, "reg_mlp." + li + ".2", 3, PE, &tw.r2_T[l], st))) return rc;
        if ((rc = p_vec(h, "reg_mlp." + li + ".2.bias", 3, &tw.r2_b[l]))) return rc;
    }
    if ((rc = h->pool.alloc(&h->d_w, 1))) return rc;
    EGR_CUDA_OK(cudaMemcpyAsync(h->d_w, &tw, sizeof(PoseTokenW), cudaMemcpyHostToDevice, st));
    h->tw_host = tw;
    h->tokb = is16(h->prec) && g_opt_tc && g_opt_tok_batched;
    if (h->tokb) {
        if ((rc = gemm_tc_init())) return rc;
        if ((rc = p_alloc_wmat(h, h->tk_g2, PE, PE, st)) || (rc = p_alloc_wmat(h, h->tk_g4, PE, PE, st))) return rc;
        if ((rc = p_copy_rows(h, "query_gen_mlp.2", h->tk_g2.f32, h->tk_g2.bias, PE, PE, st))) return rc;
        if ((rc = p_copy_rows(h, "query_gen_mlp.4", h->tk_g4.f32, h->tk_g4.bias, PE, PE, st))) return rc;
        if ((rc = round_tf32_inplace(h->tk_g2.f32, PE * PE, st)) || (rc = round_tf32_inplace(h->tk_g4.f32, PE * PE, st))) return rc;
        std::vector<const float*> ptrs(4 * 8, nullptr);
        for (int l = 0; l < h->L; ++l) {
            if ((rc = p_build_tokb_layer(h, l, tw.layer[l], st))) return rc;
            const LayerW& lw = tw.layer[l];
            const float* arr[8] = {lw.lnc_w, lw.lnc_b, lw.lns_w, lw.lns_b, lw.lnf_w, lw.lnf_b, tw.pn_w[l], tw.pn_b[l]};
            for (int i = 0; i < 8; ++i) ptrs[l * 8 + i] = arr[i];
        }
        if ((rc = h->pool.alloc(&h->d_ptrs, 4 * 8))) return rc;
        EGR_CUDA_OK(cudaMemcpyAsync(h->d_ptrs, ptrs.data(), sizeof(const float*) * 4 * 8, cudaMemcpyHostToDevice, st));
        EGR_CUDA_OK(cudaStreamSynchronize(st));
    }
    EGR_CUDA_OK(cudaStreamSynchronize(st));
    h->packed = true;
    return EGR_OK;
}

extern "C" int64_t egr_pose3d_workspace_bytes(egr_pose3d* h, int B) {
    if (!h || B <= 0) return 0;
    return p_carve(h, B, nullptr, 0, nullptr);
}

extern "C" int egr_pose3d_forward(egr_pose3d* h, int B, const float* feats_init, const float* feats_final,
                                  const float* coord_trans_mat, float* preds, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
    EGR_CHECK(h, EGR_ERR_INVALID, "pose3d_forward: null handle");
    if (int rc = require_device()) return rc;
    EGR_CHECK(h->packed, EGR_ERR_STATE, "pose3d_forward: parameters changed or never packed; call egr_pose3d_prepack");
    EGR_CHECK(B > 0 && preds && workspace, EGR_ERR_INVALID, "pose3d_forward: null pointer");
    EGR_CHECK((feats_final || h->st_final_tf32 || h->st_final_f16) && (feats_init || feats_final || h->st_sampled), EGR_ERR_INVALID,
              "pose3d_forward: a NULL NCHW input needs its staged channels-last copy (egr_pose3d_use_staged)");
    const int is_rw = (h->cam_model & 1);
    EGR_CHECK(!is_rw || coord_trans_mat, EGR_ERR_INVALID, "pose3d_forward: ego4view_rw* needs coord_trans_mat [B,V,4,4] fp32");
    PBufs w;
    const int64_t need = p_carve(h, B, workspace, workspace_bytes, &w);
    EGR_CHECK(need <= workspace_bytes, EGR_ERR_STATE, "pose3d_forward: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    const int V = h->V, J = h->J;
    const int bfs = act_code(h->prec);                     // sampled-map dtype: 0 fp32, 1 bf16, 2 fp16
    const int prec = p2_prec(h), bf = (prec == EGR_PREC_BF16);   // proposal-branch dtype
    const int VB = V * B;
    int rc;
    gemm_tc_set_scratch(w.splitk);
    EGR_MARK("P_stage_nhwc", st);
    // staging copies; the sampled map is frame_feats_init when use_pred_heatmap_init (:424-427)
    const float* sampled = h->use_init ? feats_init : feats_final;
    const int rnd = (prec == PREC_TF32);      // every operand of a tf32 stage is pre-rounded to the nearest TF32 value
    const int f16 = (prec == PREC_FP16);
    // staged copies left by a chained mvfex forward replace the staging passes (one-shot)
    const void* st_s = h->st_sampled;
    const float* st_f = h->st_final_tf32;
    const int st_s_bf16 = h->st_sampled_bf16;
    const void* st_fh = h->st_final_f16;
    h->st_sampled = nullptr; h->st_final_tf32 = nullptr; h->st_final_f16 = nullptr;
    const void* Xf = w.Xf;
    if (st_f && rnd) Xf = st_f;
    else if (st_fh && f16) Xf = st_fh;
    else {
        EGR_CHECK(feats_final, EGR_ERR_INVALID, "pose3d_forward: feats_final is NULL and its staged copy does not fit this precision");
        if ((rc = nchw_to_nhwc(feats_final, w.Xf, B, V, PC, PHW, f16 ? 3 : rnd ? 2 : bf, st))) return rc;
    }
    const void* Xs = Xf;
    const int pcode = f16 ? 2 : bf ? 1 : 0;                // dtype of Xf; it doubles as the sampled map only when both agree
    if (st_s && st_s_bf16 == bfs) Xs = st_s;
    else if (sampled != feats_final || bfs != pcode || rnd || (f16 && bfs != 2)) {
        EGR_CHECK(sampled, EGR_ERR_INVALID, "pose3d_forward: the sampled map is NULL and its staged copy does not fit this precision");
        if ((rc = nchw_to_nhwc(sampled, w.Xi, B, V, PC, PHW, bfs == 2 ? 3 : bfs, st))) return rc;
        Xs = w.Xi;
    }
    EGR_MARK("P2a", st);
    // P2 conv_frame_feat
    GemmDesc d;
    d.A = Xf; d.lda = PC; d.M = VB * PHW; d.D = w.p0; d.ldd = 64; d.epi = EPI_RELU; d.round_tf32 = rnd;
    if ((rc = run_gemm(d, h->c0, 0, prec, false, st))) return rc;
    EGR_MARK("P2b", st);
    d = GemmDesc();
    d.A = w.p0; d.amode = A_CONV3S2; d.Hin = 64; d.Win = 64; d.Cin = 64; d.M = VB * 1024; d.D = w.p2; d.ldd = 128; d.epi = EPI_RELU; d.round_tf32 = rnd;
    if ((rc = run_gemm(d, h->c2, 0, prec, false, st))) return rc;
    EGR_MARK("P2pool", st);
    if ((rc = maxpool2_nhwc(w.p2, w.p3, f16 ? 2 : bf, VB, 32, 32, 128, st))) return rc;
    EGR_MARK("P2c", st);
    d = GemmDesc();
    d.A = w.p3; d.lda = 128; d.M = VB * 256; d.D = w.p5; d.ldd = 64; d.epi = EPI_RELU; d.round_tf32 = rnd;
    if ((rc = run_gemm(d, h->c5, 0, prec, false, st))) return rc;
    EGR_MARK("P2d", st);
    d = GemmDesc();
    d.A = w.p5; d.amode = A_CONV3S2; d.Hin = 16; d.Win = 16; d.Cin = 64; d.M = VB * 64; d.D = w.p7; d.ldd = 128; d.epi = EPI_RELU; d.round_tf32 = rnd;
    if ((rc = run_gemm(d, h->c7, 0, prec, false, st))) return rc;
    EGR_MARK("P2mlp0", st);
    // mlp_pred: K-split over the V view blocks of p7 ([V][B][64*128])
    d = GemmDesc();
    d.A = w.p7; d.lda = 64 * 128; d.kblk = 64 * 128; d.kblk_stride = (int64_t)B * 64 * 128; d.M = B; d.D = w.m0; d.ldd = h->m0.N;
    d.epi = EPI_GELU; d.round_tf32 = rnd || f16;      // operand of the TF32 mlp_pred.1
    if ((rc = run_gemm(d, h->m0, 0, prec, /*out_f32=*/true, st))) return rc;
    EGR_MARK("P2mlp12", st);
    {   // Linear(2048 -> 128) GELU on the tensor cores in the tf32 branch; the last Linear(128 -> 48) stays fp32 SIMT
        GemmDesc t;
        t.A = w.m0; t.lda = h->m1.K; t.M = B; t.D = w.m1; t.ldd = h->m1.N; t.epi = EPI_GELU;
        if (rnd || f16) {
            if ((rc = run_gemm(t, h->m1, 0, PREC_TF32, true, st))) return rc;
        } else {
            t.N = h->m1.N; t.K = h->m1.K; t.W = h->m1.f32; t.bias = h->m1.bias;
            if ((rc = gemm_simt(t, 0, 0, st))) return rc;
        }
        t = GemmDesc();
        t.A = w.m1; t.lda = h->m2.K; t.M = B; t.N = h->m2.N; t.K = h->m2.K; t.W = h->m2.f32; t.bias = h->m2.bias;
        t.D = preds; t.ldd = h->m2.N; t.epi = EPI_NONE;     // preds[0] = proposal
        if ((rc = gemm_simt(t, 0, 0, st))) return rc;
    }
    EGR_MARK("P4tokens", st);
    // P3 + P4
    PoseTokenArgs ta{};
    ta.B = B; ta.V = V; ta.J = J; ta.H = PH; ta.W = PW; ta.L = h->L; ta.is_rw = is_rw;
    for (int v = 0; v < 4; ++v) { ta.cam_id[v] = h->cam_id[v]; ta.cam[v] = h->cam[v]; }
    ta.ctm = coord_trans_mat; ta.mlp_pred = preds; ta.preds = preds; ta.X = Xs; ta.w = h->d_w;
    ta.dbg_anchors = w.anch; ta.dbg_valid = w.valid;
    if (h->tokb) {
        if ((rc = p_run_tokens_batched(h, B, w, Xs, bfs, coord_trans_mat, preds, h->tw_host, st))) return rc;
    } else {
        if ((rc = launch_pose_tokens(ta, bfs == 1, st))) return rc;
    }
    EGR_MARK(nullptr, st);
    const int64_t s = (bf || f16) ? 2 : 4;
    h->dbg["p7"] = std::make_pair(w.p7, (int64_t)VB * 64 * 128 * s);
    h->dbg["p0"] = std::make_pair(w.p0, (int64_t)VB * PHW * 64 * s);
    h->dbg["m0"] = std::make_pair((void*)w.m0, (int64_t)B * 2048 * 4);
    h->dbg["anchors"] = std::make_pair((void*)w.anch, (int64_t)VB * J * 8);
    h->dbg["valid"] = std::make_pair((void*)w.valid, (int64_t)VB * J);
    return EGR_OK;
}

extern "C" int egr_pose3d_use_staged(egr_pose3d* h, const void* sampled_nhwc, int sampled_is_bf16, const float* final_nhwc_tf32) {
    EGR_CHECK(h, EGR_ERR_INVALID, "pose3d_use_staged: null handle");
    h->st_sampled = sampled_nhwc; h->st_sampled_bf16 = sampled_is_bf16; h->st_final_tf32 = final_nhwc_tf32;   // dtype code 0 fp32 / 1 bf16 / 2 fp16
    return EGR_OK;
}

extern "C" int egr_pose3d_use_staged_final_f16(egr_pose3d* h, const void* final_nhwc_f16) {
    EGR_CHECK(h, EGR_ERR_INVALID, "pose3d_use_staged_final_f16: null handle");
    h->st_final_f16 = final_nhwc_f16;
    return EGR_OK;
}

extern "C" int egr_pose3d_proposal_dtype(egr_pose3d* h) {      // 0 fp32 (SIMT), 1 bf16, 2 TF32, 3 fp16
    return h ? p2_prec(h) : -1;
}

extern "C" int egr_pose3d_debug_buffer(egr_pose3d* h, const char* name, void** ptr, int64_t* bytes) {
    EGR_CHECK(h && name && ptr && bytes, EGR_ERR_INVALID, "debug_buffer: null argument");
    auto it = h->dbg.find(name);
    EGR_CHECK(it != h->dbg.end(), EGR_ERR_INVALID, "debug_buffer: unknown buffer '%s'", name);
    *ptr = it->second.first;
    *bytes = it->second.second;
    return EGR_OK;
}
