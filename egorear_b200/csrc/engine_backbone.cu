// backbone engine: the two stereo ResnetBackbone modules of EgoPoseFormerHeatmapMVFEX
// (models/backbones/resnet.py:6-152 — torchvision ResNet18 split at strides 2/4/8/16/32 + EfficientFPN, eval mode) as
// tcgen05 implicit-GEMM stages, emitting directly what the hot path reads: the stride-4 FPN map as view-major channels-last
// 16-bit [V][B][64*64][128] (egr_mvfex_use_staged_input) and the stride-32 map as fp32 [B][V][512][8][8].
//
//   stem   7x7 s2 p3 (3->64) + BN + ReLU      im2col rows (K = 147 padded to 192) -> plain GEMM          256^2 -> 128^2
//   pool   MaxPool 3x3 s2 p1                   channels-last kernel                                        -> 64^2
//   layer1 2 BasicBlocks 64->64                3x3 s1 implicit GEMMs (one 4-D TMA box per tap), BN folded, residual add + ReLU in
//   layer2-4 (128 @32^2, 256 @16^2, 512 @8^2)  the epilogue (EPI_ADD_RELU); first conv 3x3 s2 (5-D box), downsample = its centre tap
//   FPN    lateral 1x1 + ReLU per level; top-down: fuse 1x1(cat(lat_hi, up2(lat_lo))) = W_a lat_hi + up2(W_b lat_lo) (both linear:
//          the low-resolution half is convolved BEFORE the upsample, a quarter of the pixels) + ReLU in one epilogue
//          (EPI_ADDUP_RELU), then 3x3 s1 + ReLU
// Front weights serve views 0-1, back weights views 2-3: every stage is ONE launch with 2 groups (1 for the stereo configs).
// 16-bit precisions only (bf16 / fp16 operands, fp32 accumulate in TMEM); EGR_PREC_FP32 keeps the PyTorch backbone.
#include "engine_common.cuh"
#include "backbone_ops.cuh"

namespace egr {
int g_opt_stem_fused = 1;      // option "stem_fused": 0 = im2col buffer + plain GEMM (the first version, kept as a cross-check)
int g_opt_fpn_epi = 0;         // option "fpn_epi": 1 = FPN fuse conv with the upsample-add gathered in its epilogue (slower, see forward)
}
using namespace egr;

static const char* kEnc[2] = {"heatmap_estimator_stereo_front.encoder.", "heatmap_estimator_stereo_back.encoder."};

struct BbBlock { WMat c1, c2, ds; bool has_ds = false; };

struct egr_backbone {
    int V = 4, prec = EGR_PREC_BF16, sets = 2;
    bool packed = false;
    ParamTable params;
    DevPool pool;
    WMat stem;
    BbBlock blk[4][2];
    WMat lat[4], fuse_hi[3], fuse_lo[3], fpn[3];
};

namespace {

constexpr float BN_EPS = 1e-5f;
constexpr int IMG = 256;

int fetch(egr_backbone* h, const std::string& key, int64_t numel, const float** out) {
    int rc = EGR_OK;
    *out = h->params.get(key, numel, &rc);
    return rc;
}

// one conv (+ BN when `bn` is non-empty, + conv bias when has_bias) of every weight set -> WMat [sets][N][K] with its 16-bit copy
//   kind 0: 1x1 ([N][Cin]; col0/ncol select a column range of the stored [N][Ktot] weight: the two halves of fuse_convs)
//   kind 1: 3x3 -> [N][ky][kx][Cin]      kind 2: stem 7x7 -> [N][STEM_K]
int bb_make(egr_backbone* h, WMat& m, int N, int K, int kind, const std::string& conv, const std::string& bn, bool has_bias,
            cudaStream_t st, int Ktot = 0, int col0 = 0, bool want_bias = true) {
    const int S = h->sets;
    m.N = N; m.K = K; m.sets = S;
    int rc;
    if ((rc = h->pool.alloc(&m.f32, (int64_t)S * N * K))) return rc;
    if ((rc = h->pool.alloc(&m.bias, (int64_t)S * N))) return rc;
    EGR_CUDA_OK(cudaMemsetAsync(m.bias, 0, sizeof(float) * S * N, st));
    float* tmp = nullptr;
    if ((rc = h->pool.alloc(&tmp, (int64_t)N * K))) return rc;
    for (int s = 0; s < S; ++s) {
        const std::string p = kEnc[s];
        float* dst = m.f32 + (int64_t)s * N * K;
        const float* w;
        float* packed = bn.empty() ? dst : tmp;
        if (kind == 0) {
            const int kt = Ktot ? Ktot : K;
            if ((rc = fetch(h, p + conv + ".weight", (int64_t)N * kt, &w))) return rc;
            EGR_CUDA_OK(cudaMemcpy2DAsync(packed, sizeof(float) * K, w + col0, sizeof(float) * kt, sizeof(float) * K, N,
                                          cudaMemcpyDeviceToDevice, st));
        } else if (kind == 1) {
            if ((rc = fetch(h, p + conv + ".weight", (int64_t)N * K, &w))) return rc;
            if ((rc = repack_conv3(w, packed, N, K / 9, st))) return rc;
        } else {
            if ((rc = fetch(h, p + conv + ".weight", (int64_t)N * 147, &w))) return rc;
            if ((rc = stem_repack(w, packed, N, st))) return rc;
        }
        if (!bn.empty()) {
            const float *g, *b, *mu, *var;
            if ((rc = fetch(h, p + bn + ".weight", N, &g)) || (rc = fetch(h, p + bn + ".bias", N, &b)) ||
                (rc = fetch(h, p + bn + ".running_mean", N, &mu)) || (rc = fetch(h, p + bn + ".running_var", N, &var)))
                return rc;
            if ((rc = bn_fold(tmp, g, b, mu, var, BN_EPS, dst, m.bias + (int64_t)s * N, N, K, st))) return rc;
        } else if (has_bias && want_bias) {
            const float* b;
            if ((rc = fetch(h, p + conv + ".bias", N, &b))) return rc;
            EGR_CUDA_OK(cudaMemcpyAsync(m.bias + (int64_t)s * N, b, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (h->prec == EGR_PREC_FP16) {
        if ((rc = h->pool.alloc(&m.f16, (int64_t)S * N * K))) return rc;
        if ((rc = cast_f16(m.f32, m.f16, (int64_t)S * N * K, st))) return rc;
    } else {
        if ((rc = h->pool.alloc(&m.bf16, (int64_t)S * N * K))) return rc;
        if ((rc = cast_bf16(m.f32, m.bf16, (int64_t)S * N * K, st))) return rc;
    }
    return EGR_OK;
}

const int kCh[4] = {64, 128, 256, 512};
const int kRes[4] = {64, 32, 16, 8};
const char* kLayer[4] = {"backbone.layer_s4.1", "backbone.layer_s8", "backbone.layer_s16", "backbone.layer_s32"};

struct BbBufs {
    void *col, *s2, *x[4], *t[4], *y[4], *idn[4], *lat[4], *low[3], *f[3], *g[2];
};

int64_t bb_carve(const egr_backbone* h, int B, void* base, int64_t cap, BbBufs* o) {
    const int64_t VB = (int64_t)h->V * B, e = 2;
    Carver c(base, cap);
    BbBufs b{};
    b.col = g_opt_stem_fused ? nullptr : c.take(VB * 128 * 128 * STEM_K * e);
    b.s2 = c.take(VB * 128 * 128 * 64 * e);
    for (int l = 0; l < 4; ++l) {
        const int64_t n = VB * kRes[l] * kRes[l] * kCh[l] * e;
        b.x[l] = c.take(n); b.t[l] = c.take(n); b.y[l] = c.take(n); b.idn[l] = c.take(n);
        b.lat[l] = c.take(VB * kRes[l] * kRes[l] * 128 * e);
    }
    for (int i = 0; i < 3; ++i) {
        b.low[i] = c.take(VB * kRes[i + 1] * kRes[i + 1] * 128 * e);      // W_b lat[i+1] at the low resolution
        b.f[i] = c.take(VB * kRes[i] * kRes[i] * 128 * e);                // fused map at level i
    }
    b.g[0] = c.take(VB * kRes[2] * kRes[2] * 128 * e);                    // fpn_convs outputs of levels 2 and 1 (level 0 = the result)
    b.g[1] = c.take(VB * kRes[1] * kRes[1] * 128 * e);
    if (o) *o = b;
    return c.off + 256;
}

}  // namespace

extern "C" int egr_backbone_create(int num_views, int precision, egr_backbone** out) {
    EGR_CHECK(out, EGR_ERR_INVALID, "backbone_create: null out");
    EGR_CHECK(num_views == 2 || num_views == 4, EGR_ERR_UNSUPPORTED, "backbone: num_views=%d", num_views);
    EGR_CHECK(precision == EGR_PREC_BF16 || precision == EGR_PREC_FP16, EGR_ERR_UNSUPPORTED,
              "backbone: tensor-core precisions only (EGR_PREC_BF16 / EGR_PREC_FP16); EGR_PREC_FP32 keeps the PyTorch backbone");
    if (int rc = require_device()) return rc;
    egr_backbone* h = new egr_backbone();
    h->V = num_views; h->prec = precision; h->sets = num_views / 2;
    *out = h;
    return EGR_OK;
}

extern "C" int egr_backbone_destroy(egr_backbone* h) {
    if (!h) return EGR_OK;
    cudaDeviceSynchronize();
    h->pool.release();
    delete h;
    return EGR_OK;
}

extern "C" int egr_backbone_set_param(egr_backbone* h, const char* key, const float* ptr, int64_t numel) {
    EGR_CHECK(h, EGR_ERR_INVALID, "backbone_set_param: null handle");
    h->packed = false;
    return h->params.set(key, ptr, numel);
}

extern "C" int egr_backbone_prepack(egr_backbone* h, void* stream) {
    EGR_CHECK(h, EGR_ERR_INVALID, "backbone_prepack: null handle");
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    EGR_CUDA_OK(cudaStreamSynchronize(st));
    h->pool.release();
    h->packed = false;
    int rc;
    if ((rc = gemm_tc_init())) return rc;
    if ((rc = bb_make(h, h->stem, 64, STEM_K, 2, "backbone.layer_s2.0", "backbone.layer_s2.1", false, st))) return rc;
    for (int l = 0; l < 4; ++l) {
        const int C = kCh[l], Cin = l ? kCh[l - 1] : 64;
        for (int b = 0; b < 2; ++b) {
            const std::string p = std::string(kLayer[l]) + "." + std::to_string(b);
            BbBlock& B = h->blk[l][b];
            const int cin1 = (b == 0) ? Cin : C;
            if ((rc = bb_make(h, B.c1, C, 9 * cin1, 1, p + ".conv1", p + ".bn1", false, st))) return rc;
            if ((rc = bb_make(h, B.c2, C, 9 * C, 1, p + ".conv2", p + ".bn2", false, st))) return rc;
            B.has_ds = (b == 0 && l > 0);
            if (B.has_ds && (rc = bb_make(h, B.ds, C, Cin, 0, p + ".downsample.0", p + ".downsample.1", false, st))) return rc;
        }
    }
    for (int i = 0; i < 4; ++i)
        if ((rc = bb_make(h, h->lat[i], 128, kCh[i], 0, "neck.lateral_convs." + std::to_string(i) + ".0", "", true, st))) return rc;
    for (int i = 0; i < 3; ++i) {
        const std::string f = "neck.fuse_convs." + std::to_string(i) + ".0";
        // cat((lat_hi, up(lat_lo)), 1): input channels [0,128) multiply the same-resolution map, [128,256) the upsampled one
        if ((rc = bb_make(h, h->fuse_hi[i], 128, 128, 0, f, "", true, st, 256, 0, true))) return rc;
        if ((rc = bb_make(h, h->fuse_lo[i], 128, 128, 0, f, "", true, st, 256, 128, false))) return rc;
        if ((rc = bb_make(h, h->fpn[i], 128, 9 * 128, 1, "neck.fpn_convs." + std::to_string(i) + ".0", "", true, st))) return rc;
    }
    EGR_CUDA_OK(cudaStreamSynchronize(st));
    h->packed = true;
    return EGR_OK;
}

extern "C" int64_t egr_backbone_workspace_bytes(egr_backbone* h, int B) {
    if (!h || B <= 0) return 0;
    return bb_carve(h, B, nullptr, 0, nullptr);
}

extern "C" int egr_backbone_forward(egr_backbone* h, int B, const float* img, void* feat_staged, float* bfb, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
    EGR_CHECK(h, EGR_ERR_INVALID, "backbone_forward: null handle");
    if (int rc = require_device()) return rc;
    EGR_CHECK(h->packed, EGR_ERR_STATE, "backbone_forward: parameters changed or never packed; call egr_backbone_prepack");
    EGR_CHECK(B > 0 && img && feat_staged && bfb && workspace, EGR_ERR_INVALID, "backbone_forward: null pointer / empty batch");
    EGR_CHECK(((uintptr_t)feat_staged & 15) == 0 && ((uintptr_t)img & 15) == 0, EGR_ERR_INVALID, "backbone_forward: pointers must be 16-byte aligned");
    BbBufs w;
    const int64_t need = bb_carve(h, B, workspace, workspace_bytes, &w);
    EGR_CHECK(need <= workspace_bytes, EGR_ERR_STATE, "backbone_forward: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    const int V = h->V, G = h->sets, vpg = V / G, prec = h->prec, dt = act_code(prec);
    const int64_t IMGS = (int64_t)vpg * B;          // images per group
    int rc;
    gemm_tc_set_scratch(nullptr);
    auto conv = [&](const void* A, int amode, int Hin, int Cin, const WMat& W, void* D, int epi, const void* aux, int Hout) -> int {
        GemmDesc d;
        const int64_t M = IMGS * Hout * Hout;
        EGR_CHECK(M < (1ll << 31), EGR_ERR_UNSUPPORTED, "backbone_forward: %lld rows per group exceed the 32-bit tile index; split the batch", (long long)M);
        d.A = A; d.amode = amode; d.M = (int)M; d.D = D; d.ldd = W.N; d.epi = epi; d.aux = aux;
        d.groups = G; d.d_gs = M * W.N;
        if (amode == A_PLAIN) { d.lda = Cin; d.a_gs = M * Cin; }
        else { d.Hin = Hin; d.Win = Hin; d.Cin = Cin; d.a_gs = IMGS * Hin * Hin * Cin; }
        if (epi == EPI_ADD_RELU) d.aux_gs = M * W.N;
        if (epi == EPI_ADDUP_RELU) { d.Hout = Hout; d.Wout = Hout; d.aux_gs = IMGS * (Hout / 2) * (Hout / 2) * W.N; }
        d.ws = 1;       // weight-stationary wherever the slab fits beside >= 4 A stages (layer1's N = 64 convs, the 1x1 convs)
        return run_gemm(d, W, 0, prec, false, st);
    };
    EGR_MARK("B_stem", st);
    if (g_opt_stem_fused && (int64_t)V * B <= 65535) {
        // one tensor-core kernel: im2col rows are built in shared memory (stem_tc.cu)
        const void* w16 = (prec == EGR_PREC_FP16) ? (const void*)h->stem.f16 : (const void*)h->stem.bf16;
        if ((rc = stem_tc(img, w16, h->stem.bias, dt, B, V, vpg, w.s2, st))) return rc;
    } else {
        if ((rc = stem_im2col(img, w.col, dt, B, V, IMG, IMG, st))) return rc;
        if ((rc = conv(w.col, A_PLAIN, 0, STEM_K, h->stem, w.s2, EPI_RELU, nullptr, 128))) return rc;
    }
    if ((rc = maxpool3s2_nhwc(w.s2, w.x[0], dt, (int64_t)V * B, 128, 128, 64, st))) return rc;
    const void* x = w.x[0];       // input of the current layer
    for (int l = 0; l < 4; ++l) {
        static const char* names[4] = {"B_layer1", "B_layer2", "B_layer3", "B_layer4"};
        EGR_MARK(names[l], st);
        const int C = kCh[l], R = kRes[l], Cin = l ? kCh[l - 1] : 64, Rin = l ? kRes[l - 1] : 64;
        // block 0
        const BbBlock& b0 = h->blk[l][0];
        const void* identity = x;
        if (b0.has_ds) {
            if ((rc = conv(x, A_CONV3S2, Rin, Cin, b0.c1, w.t[l], EPI_RELU, nullptr, R))) return rc;
            if ((rc = conv(x, A_CONV3S2, Rin, Cin, b0.ds, w.idn[l], EPI_NONE, nullptr, R))) return rc;      // K == Cin: centre tap
            identity = w.idn[l];
        } else if ((rc = conv(x, A_CONV3S1, R, Cin, b0.c1, w.t[l], EPI_RELU, nullptr, R))) return rc;
        if ((rc = conv(w.t[l], A_CONV3S1, R, C, b0.c2, w.y[l], EPI_ADD_RELU, identity, R))) return rc;
        // block 1
        const BbBlock& b1 = h->blk[l][1];
        if ((rc = conv(w.y[l], A_CONV3S1, R, C, b1.c1, w.t[l], EPI_RELU, nullptr, R))) return rc;
        void* out = (l == 0) ? w.idn[0] : w.x[l];      // a buffer of this level that is free here (layer 1 has no downsample branch)
        if ((rc = conv(w.t[l], A_CONV3S1, R, C, b1.c2, out, EPI_ADD_RELU, w.y[l], R))) return rc;
        x = out;
        // lateral 1x1 + ReLU of this level
        if ((rc = conv(x, A_PLAIN, 0, C, h->lat[l], w.lat[l], EPI_RELU, nullptr, R))) return rc;
    }
    EGR_MARK("B_bottom", st);
    if ((rc = bottom_to_nchw(x, bfb, dt, B, V, 64, 512, st))) return rc;
    EGR_MARK("B_fpn", st);
    const void* top = w.lat[3];
    for (int i = 2; i >= 0; --i) {
        const int R = kRes[i];
        // low-resolution half of the fuse conv, BEFORE the upsample (no bias, no activation)
        if ((rc = conv(top, A_PLAIN, 0, 128, h->fuse_lo[i], w.low[i], EPI_NONE, nullptr, kRes[i + 1]))) return rc;
        if (g_opt_fpn_epi) {
            // one kernel: the upsampled half is gathered in the GEMM epilogue (row-per-lane gathers: L1-wavefront bound, measured
            // 1.02 ms at 64^2 for 128 frames vs 0.27 + 0.2 ms for the two-kernel form below)
            if ((rc = conv(w.lat[i], A_PLAIN, 0, 128, h->fuse_hi[i], w.f[i], EPI_ADDUP_RELU, w.low[i], R))) return rc;
        } else {
            if ((rc = conv(w.lat[i], A_PLAIN, 0, 128, h->fuse_hi[i], w.f[i], EPI_NONE, nullptr, R))) return rc;
            if ((rc = up2_add_relu_nhwc(w.f[i], w.low[i], w.f[i], dt, (int64_t)V * B, R, 128, st))) return rc;
        }
        void* out = (i == 2) ? w.g[0] : (i == 1) ? w.g[1] : feat_staged;      // level 0 IS the staged input of the hot path
        if ((rc = conv(w.f[i], A_CONV3S1, R, 128, h->fpn[i], out, EPI_RELU, nullptr, R))) return rc;
        top = out;
    }
    EGR_MARK(nullptr, st);
    return EGR_OK;
}
