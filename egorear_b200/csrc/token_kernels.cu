// The two fused token kernels: one CTA per (frame, refiner) for the mvfex jqa layer, one CTA per frame for the
// pose3d lifting transformer.  All math fp32; features are gathered from the channels-last staging copy.
#include "token_kernels.cuh"

namespace egr {

// =====================================================================================================
// mvfex: Q1 (jqa query) -> A1 A2 A3 -> post_norm -> transposed token image for the T1 GEMMs
// =====================================================================================================
constexpr int MV_E = 256, MV_J = 15;

template <typename T>
__global__ void __launch_bounds__(MV_E, 1)
mvf_token_kernel(MvfTokenArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                                   // [J][E]
    float* ys = xs + MV_J * MV_E;                       // [J][E]
    float* big = ys + MV_J * MV_E;                      // token_big_floats
    float* probs = big + token_big_floats<MV_E, MV_J>();   // [NH*J*16]
    float* s_anch = probs + TOK_NH * MV_J * 16;         // [V][J][2]
    __shared__ uint8_t s_valid[4 * MV_J];
    __shared__ const T* s_X[4];

    const int b = blockIdx.x;
    const int g = blockIdx.y;                           // data group
    const MvfTokenW& w = a.w[a.r0 + g];                 // refiner weights
    const int tid = threadIdx.x, n = tid;
    constexpr int E = MV_E, J = MV_J;

    for (int i = tid; i < a.V * J * 2; i += E) {
        const int v = i / (J * 2), rr = i - v * J * 2;
        s_anch[i] = a.anchors[((int64_t)b * a.V + v) * J * 2 + rr];
    }
    for (int i = tid; i < a.V * J; i += E) {
        const int v = i / J, j = i - v * J;
        s_valid[i] = a.valid[((int64_t)b * a.V + v) * J + j];
    }
    if (tid < a.V) s_X[tid] = reinterpret_cast<const T*>(a.X) + ((int64_t)tid * a.B + b) * a.H * a.W * TOK_RAWC;

    // ---- Q1: x = relu(fc_query(embed + fc_bfb(avgpool(bfb)) + heatmap_proj.2(q1))) ----
    const float* q1 = a.q1 + ((int64_t)g * a.B + b) * J * E;
    for (int i = tid; i < J * E; i += E) ys[i] = q1[i];
    // adaptive_avg_pool2d(bfb, 1): 512 channels, thread handles channels tid and tid+256
    {
        const float* bf = a.bfb + (int64_t)g * a.bfb_gs + (int64_t)b * a.bfb_bs;
        for (int c = tid; c < 512; c += E) {
            const float* p = bf + (int64_t)c * a.bfb_hw;
            float s = 0.f;
            for (int i = 0; i < a.bfb_hw; ++i) s += p[i];
            big[c] = s / (float)a.bfb_hw;
        }
    }
    __syncthreads();
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.hp2_T, E, E, n, ys, E, acc);
        float g = __ldg(w.bfb_b + n);
        for (int k = 0; k < 512; ++k) g = fmaf(__ldg(w.bfb_T + (int64_t)k * E + n), big[k], g);
        const float hb = __ldg(w.hp2_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) xs[j * E + n] = __ldg(w.jq + j * E + n) + g + (acc[j] + hb);
    }
    __syncthreads();
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.fcq_T, E, E, n, xs, E, acc);
        const float qb = __ldg(w.fcq_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) ys[j * E + n] = fmaxf(acc[j] + qb, 0.f);
    }
    __syncthreads();
    for (int i = tid; i < J * E; i += E) xs[i] = ys[i];
    __syncthreads();

    token_layer<E, J, T, true>(w.layer, a.V, a.H, a.W, s_X, s_anch, s_valid, xs, ys, big, probs);

    // ---- post_norm, then write the [J][16x16] token image transposed: xT[pos][j], j padded to 16 ----
    layernorm_rows<E>(xs, ys, E, E, J, w.pn_w, w.pn_b);
    __syncthreads();
    T* out = reinterpret_cast<T*>(a.xT) + ((int64_t)g * a.B + b) * E * 16;
    for (int i = tid; i < E * 16; i += E) {
        const int pos = i >> 4, j = i & 15;
        ActT<T>::st(out + i, (j < J) ? ys[j * E + pos] : 0.f);
    }
}

int launch_mvf_tokens(const MvfTokenArgs& a, int act_bf16, cudaStream_t st) {
    EGR_CHECK(a.J == MV_J && a.V <= 4, EGR_ERR_UNSUPPORTED, "mvf tokens: J=%d V=%d", a.J, a.V);
    const size_t smem = sizeof(float) * (2 * MV_J * MV_E + token_big_floats<MV_E, MV_J>() + TOK_NH * MV_J * 16 + 4 * MV_J * 2);
    dim3 grid(a.B, a.G);
    if (act_bf16) {
        auto k = mvf_token_kernel<__nv_bfloat16>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, MV_E, smem, st>>>(a);
    } else {
        auto k = mvf_token_kernel<float>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, MV_E, smem, st>>>(a);
    }
    EGR_LAUNCHED();
    return EGR_OK;
}

// =====================================================================================================
// pose3d: P3 reprojection -> query_gen_mlp -> L x (A1 A2 A3) -> post_norm + reg_mlp -> preds
// =====================================================================================================
constexpr int PS_E = 128, PS_J = 16;

template <typename T>
__global__ void __launch_bounds__(PS_E, 2)
pose_token_kernel(PoseTokenArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;
    float* ys = xs + PS_J * PS_E;
    float* big = ys + PS_J * PS_E;
    float* probs = big + token_big_floats<PS_E, PS_J>();
    float* s_anch = probs + TOK_NH * PS_J * 16;          // [V][J][2]
    float* s_p3 = s_anch + 4 * PS_J * 2;                 // [J][4]: (j+1)/J, x, y, z  (anchors after the in-place quirk)
    __shared__ uint8_t s_valid[4 * PS_J];
    __shared__ const T* s_X[4];

    constexpr int E = PS_E, J = PS_J;
    const int b = blockIdx.x, tid = threadIdx.x, n = tid;
    const PoseTokenW& w = *a.w;
    if (tid < a.V) s_X[tid] = reinterpret_cast<const T*>(a.X) + ((int64_t)tid * a.B + b) * a.H * a.W * TOK_RAWC;

    // ---- P3: reprojection; thread j owns joint j and walks the cameras in concat order ----
    if (tid < J) {
        const int j = tid;
        float px = a.mlp_pred[((int64_t)b * J + j) * 3 + 0];
        float py = a.mlp_pred[((int64_t)b * J + j) * 3 + 1];
        float pz = a.mlp_pred[((int64_t)b * J + j) * 3 + 2];
        for (int v = 0; v < a.V; ++v) {
            float cx_, cy_, cz_;
            if (a.is_rw) {
                const float* M = a.ctm + ((int64_t)b * a.V + v) * 16;
                const float hx = px * 0.01f, hy = py * 0.01f, hz = pz * 0.01f;
                cx_ = (M[0] * hx + M[1] * hy + M[2] * hz + M[3]) * 100.f;
                cy_ = (M[4] * hx + M[5] * hy + M[6] * hz + M[7]) * 100.f;
                cz_ = (M[8] * hx + M[9] * hy + M[10] * hz + M[11]) * 100.f;
            } else {
                float ox, oy;
                bool flip;
                syn_offset(a.cam_id[v], ox, oy, flip);
                if (flip) { px = -px; py = -py; }
                px = __fadd_rn(px, ox);
                py = __fadd_rn(py, oy);
                cx_ = px; cy_ = py; cz_ = pz;
            }
            float u, vv;
            bool fov;
            fisheye_project(cx_, cy_, cz_, a.cam[v], u, vv, fov);
            s_anch[(v * J + j) * 2 + 0] = u;
            s_anch[(v * J + j) * 2 + 1] = vv;
            s_valid[v * J + j] = fov ? 1 : 0;
            if (a.dbg_anchors) {
                a.dbg_anchors[(((int64_t)b * a.V + v) * J + j) * 2 + 0] = u;
                a.dbg_anchors[(((int64_t)b * a.V + v) * J + j) * 2 + 1] = vv;
            }
            if (a.dbg_valid) a.dbg_valid[((int64_t)b * a.V + v) * J + j] = fov ? 1 : 0;
        }
        s_p3[j * 4 + 0] = (float)(j + 1) / (float)J;
        s_p3[j * 4 + 1] = px; s_p3[j * 4 + 2] = py; s_p3[j * 4 + 3] = pz;
    }
    __syncthreads();

    // ---- query_gen_mlp: Linear(4,E) ReLU Linear(E,E) ReLU Linear(E,E) ----
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.g0_T, 4, E, n, s_p3, 4, acc);
        const float b0 = __ldg(w.g0_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) xs[j * E + n] = fmaxf(acc[j] + b0, 0.f);
    }
    __syncthreads();
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.g2_T, E, E, n, xs, E, acc);
        const float b2 = __ldg(w.g2_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) ys[j * E + n] = fmaxf(acc[j] + b2, 0.f);
    }
    __syncthreads();
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.g4_T, E, E, n, ys, E, acc);
        const float b4 = __ldg(w.g4_b + n);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < J; ++j) xs[j * E + n] = acc[j] + b4;
    }
    __syncthreads();

    for (int l = 0; l < a.L; ++l) {
        token_layer<E, J, T, false>(w.layer[l], a.V, a.H, a.W, s_X, s_anch, s_valid, xs, ys, big, probs);
        // post_norm[l] -> reg_mlp[l]: Linear(E,E) GELU Linear(E,3); pred = offset + anchors (after the quirk)
        layernorm_rows<E>(xs, ys, E, E, J, w.pn_w[l], w.pn_b[l]);
        __syncthreads();
        {
            float acc[J];
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = 0.f;
            lin_tok<J>(w.r0_T[l], E, E, n, ys, E, acc);
            const float b0 = __ldg(w.r0_b[l] + n);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float x = acc[j] + b0;
                big[j * E + n] = 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
            }
        }
        __syncthreads();
        if (tid < J * 3) {
            const int j = tid / 3, c = tid - j * 3;
            float acc = __ldg(w.r2_b[l] + c);
            for (int k = 0; k < E; ++k) acc = fmaf(__ldg(w.r2_T[l] + k * 3 + c), big[j * E + k], acc);
            a.preds[(((int64_t)(l + 1) * a.B + b) * J + j) * 3 + c] = acc + s_p3[j * 4 + 1 + c];
        }
        __syncthreads();
    }
}

int launch_pose_tokens(const PoseTokenArgs& a, int act_bf16, cudaStream_t st) {
    EGR_CHECK(a.J == PS_J && a.V <= 4 && a.L <= 4, EGR_ERR_UNSUPPORTED, "pose tokens: J=%d V=%d L=%d", a.J, a.V, a.L);
    const size_t smem = sizeof(float) * (2 * PS_J * PS_E + token_big_floats<PS_E, PS_J>() + TOK_NH * PS_J * 16 +
                                         4 * PS_J * 2 + PS_J * 4);
    if (act_bf16) {
        auto k = pose_token_kernel<__nv_bfloat16>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<a.B, PS_E, smem, st>>>(a);
    } else {
        auto k = pose_token_kernel<float>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<a.B, PS_E, smem, st>>>(a);
    }
    EGR_LAUNCHED();
    return EGR_OK;
}

}  // namespace egr
