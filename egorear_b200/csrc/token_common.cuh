// Device helpers shared by the two fused "token" kernels (mvfex refiner layer, pose3d lifting layers).
// Tokens are few (15 / 16 joints), so one CTA owns one (frame, refiner) and keeps the token matrix in
// shared memory; thread n owns output channel n of every Linear, weights are pre-transposed to [K][N] so
// that the per-k weight read is one coalesced line, and the token values are smem broadcasts.
#pragma once
#include "common.cuh"

namespace egr {

// acc[j] += sum_k Wt[k][n] * xs[j][k]         (Wt row stride = N, xs row stride = ldx, ldx % 4 == 0, K % 4 == 0)
template <int J>
__device__ __forceinline__ void lin_tok(const float* __restrict__ Wt, int K, int N, int n, const float* xs, int ldx,
                                        float (&acc)[J]) {
    const float* w = Wt + n;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        const float w0 = __ldg(w + (int64_t)(k + 0) * N);
        const float w1 = __ldg(w + (int64_t)(k + 1) * N);
        const float w2 = __ldg(w + (int64_t)(k + 2) * N);
        const float w3 = __ldg(w + (int64_t)(k + 3) * N);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 x = *reinterpret_cast<const float4*>(xs + j * ldx + k);
            acc[j] = fmaf(w0, x.x, acc[j]);
            acc[j] = fmaf(w1, x.y, acc[j]);
            acc[j] = fmaf(w2, x.z, acc[j]);
            acc[j] = fmaf(w3, x.w, acc[j]);
        }
    }
}

// In-place LayerNorm (eps 1e-5, biased variance) of J rows of width E held in smem (row stride ld);
// out may alias in.  Warp w normalises rows w, w+nwarps, ...  Caller syncs before and after.
template <int E>
__device__ __forceinline__ void layernorm_rows(const float* in, float* out, int ld_in, int ld_out, int J,
                                               const float* __restrict__ gamma, const float* __restrict__ beta) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = warp; j < J; j += nw) {
        float v[E / 32];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < E / 32; ++i) { v[i] = in[j * ld_in + lane + 32 * i]; s += v[i]; }
        const float mean = warp_sum(s) * (1.f / E);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < E / 32; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / E) + 1e-5f);
#pragma unroll
        for (int i = 0; i < E / 32; ++i) {
            const int c = lane + 32 * i;
            out[j * ld_out + c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        }
    }
}

// One bilinear corner set of mmcv's ms_deform_attn kernel for location (loc_x, loc_y) in [0,1] on an H x W map:
// pixel = loc * size - 0.5, zero outside, point skipped unless -1 < pixel < size.
struct Corners {
    int idx[4];     // flat y*W+x of the 4 corners (valid only where w != 0)
    float w[4];     // bilinear weights, 0 for out-of-map corners
    bool any;
};
__device__ __forceinline__ Corners msda_corners(float loc_x, float loc_y, int H, int W) {
    Corners c;
    const float h_im = loc_y * (float)H - 0.5f;
    const float w_im = loc_x * (float)W - 0.5f;
    c.any = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
    const int h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
    const float hh = 1.f - lh, hw = 1.f - lw;
    const bool y0 = h_low >= 0, y1 = h_high <= H - 1, x0 = w_low >= 0, x1 = w_high <= W - 1;
    c.w[0] = (c.any && y0 && x0) ? hh * hw : 0.f;
    c.w[1] = (c.any && y0 && x1) ? hh * lw : 0.f;
    c.w[2] = (c.any && y1 && x0) ? lh * hw : 0.f;
    c.w[3] = (c.any && y1 && x1) ? lh * lw : 0.f;
    const int yl = min(max(h_low, 0), H - 1), yh = min(max(h_high, 0), H - 1);
    const int xl = min(max(w_low, 0), W - 1), xh = min(max(w_high, 0), W - 1);
    c.idx[0] = yl * W + xl; c.idx[1] = yl * W + xh; c.idx[2] = yh * W + xl; c.idx[3] = yh * W + xh;
    return c;
}

// 4 consecutive channels of a channels-last feature row
template <typename T> __device__ __forceinline__ float4 ldg4(const T* p);
template <> __device__ __forceinline__ float4 ldg4<float>(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 ldg4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

// softmax over `n` contiguous values in smem, in place (one thread)
__device__ __forceinline__ void softmax_inplace(float* v, int n) {
    float m = v[0];
    for (int i = 1; i < n; ++i) m = fmaxf(m, v[i]);
    float s = 0.f;
    for (int i = 0; i < n; ++i) { const float e = expf(v[i] - m); v[i] = e; s += e; }
    const float inv = 1.f / s;
    for (int i = 0; i < n; ++i) v[i] *= inv;
}


// -----------------------------------------------------------------------------------------------------
// One transformer layer over the joint tokens of one frame:
//   A1 deformable cross-attention over V views (+ per-view validity mask, fuse_mlp, residual, LN)
//   A2 joint self-attention (4 heads) + residual + LN
//   A3 FFN (Linear, exact GELU, Linear) + residual + LN
// Reference: MultiViewTransformerLayer (estimator/egoposeformer_heatmap_mvf_ex.py:820-935) and
// EgoPoseFormerTransformerLayer (estimator/egoposeformer_mvf_ex.py:501-588), MSDeformAttn
// (models/utils/deform_attn.py:90-168).
//
// The memory projection is folded: value = Wv·(Wp·x + bp + pos) + bv is linear in x, so
//   sum_p a_p · bilinear(value)(loc_p) = (Wv·Wp) · S + E,   S = sum_p a_p · bilinear(x)(loc_p)   [raw 128 ch]
// with E = sum_p a_p · bilinear(P')(loc_p), P' = Wv·(pos + bp) + bv (mvfex: a [V][HW][E] table, sampled with the
// same zero padding), or E = (Wv·bf + bv) · (sum of in-map bilinear weight mass) when there is no pos-embed (pose3d).
// -----------------------------------------------------------------------------------------------------
struct LayerW {
    const float *so_T, *so_b;       // sampling_offsets  [E][NH*P*2]
    const float *aw_T, *aw_b;       // attention_weights [E][NH*P]
    const float *mfold_T;           // (Wv·Wp)^T         [RAWC][E]
    const float *ptab;              // P' table          [V][HW][E]   (mvfex) or null
    const float *bfold;             // Wv·bf + bv        [E]          (pose3d) or null
    const float *op_T, *op_b;       // output_proj       [E][E]
    const float *fuse_T, *fuse_b;   // fuse_mlp          [V*E][E]
    const float *lnc_w, *lnc_b;
    const float *q_T, *q_b, *k_T, *k_b, *v_T, *v_b, *o_T, *o_b;
    const float *lns_w, *lns_b;
    const float *f1_T, *f1_b;       // [E][FF]
    const float *f2_T, *f2_b;       // [FF][E]
    const float *lnf_w, *lnf_b;
};

constexpr int TOK_NH = 4, TOK_P = 16, TOK_RAWC = 128, TOK_FF = 512;

template <int E, int J>
__host__ __device__ constexpr int token_big_floats() {
    // offs + aws + S + extra   vs   q,k,v (padded)   vs   ffn hidden
    constexpr int a = J * TOK_NH * TOK_P * 3 + J * TOK_NH * TOK_RAWC + J * E;
    constexpr int b = 3 * J * (E + 1);
    constexpr int c = J * TOK_FF;
    return (a > b ? (a > c ? a : c) : (b > c ? b : c));
}

// xs: [J][E] layer input (overwritten with the layer output); ys: [J][E] scratch; big: token_big_floats();
// probs: [NH*J*16]; anchors: smem [V][J][2]; valid: smem [V][J]; Xv[v]: this frame's [H*W][RAWC] map of view v.
template <int E, int J, typename T, bool HAS_PTAB>
__device__ void token_layer(const LayerW& w, int V, int H, int W, const T* const* Xv, const float* anchors,
                            const uint8_t* valid, float* xs, float* ys, float* big, float* probs) {
    constexpr int NH = TOK_NH, P = TOK_P, HD = E / NH, RAWC = TOK_RAWC, FF = TOK_FF;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = E / 32;
    const int n = tid;
    float* offs = big;                      // [J][NH*P*2]
    float* aws = offs + J * NH * P * 2;     // [J][NH*P]
    float* S = aws + J * NH * P;            // [J][NH][RAWC]   (later reused as o[J][E])
    float* extra = S + J * NH * RAWC;       // HAS_PTAB ? [J][NH][HD] : [J][NH]

    // ---- A1: offsets and attention logits ----
    for (int m = tid; m < NH * P * 3; m += E) {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        if (m < NH * P * 2) {
            lin_tok<J>(w.so_T, E, NH * P * 2, m, xs, E, acc);
            const float b = __ldg(w.so_b + m);
#pragma unroll
            for (int j = 0; j < J; ++j) offs[j * NH * P * 2 + m] = acc[j] + b;
        } else {
            const int mm = m - NH * P * 2;
            lin_tok<J>(w.aw_T, E, NH * P, mm, xs, E, acc);
            const float b = __ldg(w.aw_b + mm);
#pragma unroll
            for (int j = 0; j < J; ++j) aws[j * NH * P + mm] = acc[j] + b;
        }
    }
    __syncthreads();
    for (int g = tid; g < J * NH; g += E) softmax_inplace(aws + g * P, P);
    __syncthreads();

    float facc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) facc[j] = 0.f;
    const float inv_w = (float)W, inv_h = (float)H;   // offsets are divided by (W, H): deform_attn.py:133-139

    for (int v = 0; v < V; ++v) {
        // ---- sampling: warp per (joint, head), lane = 4 raw channels ----
        const T* Xb = Xv[v] + lane * 4;
        const float* ptab_v = HAS_PTAB ? (w.ptab + (int64_t)v * H * W * E) : nullptr;
        for (int pair = warp; pair < J * NH; pair += NW) {
            const int j = pair / NH, h = pair - j * NH;
            const float ax = anchors[(v * J + j) * 2 + 0], ay = anchors[(v * J + j) * 2 + 1];
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float2 e2 = make_float2(0.f, 0.f);
            float wsum = 0.f;
            const float* po = offs + j * NH * P * 2 + h * P * 2;
            const float* pa = aws + (j * NH + h) * P;
#pragma unroll 4
            for (int p = 0; p < P; ++p) {
                const float lx = ax + po[p * 2 + 0] / inv_w;
                const float ly = ay + po[p * 2 + 1] / inv_h;
                const float a = pa[p];
                const Corners c = msda_corners(lx, ly, H, W);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float coef = a * c.w[q];
                    const float4 x = ldg4<T>(Xb + (int64_t)c.idx[q] * RAWC);
                    s4.x = fmaf(coef, x.x, s4.x); s4.y = fmaf(coef, x.y, s4.y);
                    s4.z = fmaf(coef, x.z, s4.z); s4.w = fmaf(coef, x.w, s4.w);
                    if (HAS_PTAB) {
                        const float2 pv = __ldg(reinterpret_cast<const float2*>(ptab_v + (int64_t)c.idx[q] * E + h * HD) + lane);
                        e2.x = fmaf(coef, pv.x, e2.x); e2.y = fmaf(coef, pv.y, e2.y);
                    } else {
                        wsum += coef;
                    }
                }
            }
            *reinterpret_cast<float4*>(S + (j * NH + h) * RAWC + lane * 4) = s4;
            if (HAS_PTAB) *reinterpret_cast<float2*>(extra + (j * NH + h) * HD + lane * 2) = e2;
            else if (lane == 0) extra[j * NH + h] = wsum;
        }
        __syncthreads();
        // ---- value = (Wv·Wp)·S + E ----
        {
            const int h = n / HD;
            float acc[J];
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = 0.f;
            lin_tok<J>(w.mfold_T, RAWC, E, n, S + h * RAWC, NH * RAWC, acc);
            const float bf = HAS_PTAB ? 0.f : __ldg(w.bfold + n);
#pragma unroll
            for (int j = 0; j < J; ++j)
                ys[j * E + n] = acc[j] + (HAS_PTAB ? extra[(j * NH + h) * HD + (n - h * HD)] : bf * extra[j * NH + h]);
        }
        __syncthreads();
        // ---- output_proj, then zero the rows whose anchor is invalid in this view (:910 / :563) ----
        {
            float acc[J];
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = 0.f;
            lin_tok<J>(w.op_T, E, E, n, ys, E, acc);
            const float b = __ldg(w.op_b + n);
#pragma unroll
            for (int j = 0; j < J; ++j) S[j * E + n] = valid[v * J + j] ? (acc[j] + b) : 0.f;
        }
        __syncthreads();
        // ---- fuse_mlp slice of this view ----
        lin_tok<J>(w.fuse_T + (int64_t)v * E * E, E, E, n, S, E, facc);
        __syncthreads();
    }
    {
        const float b = __ldg(w.fuse_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) ys[j * E + n] = xs[j * E + n] + facc[j] + b;
    }
    __syncthreads();
    layernorm_rows<E>(ys, xs, E, E, J, w.lnc_w, w.lnc_b);
    __syncthreads();

    // ---- A2: joint self-attention ----
    constexpr int LDQ = E + 1;
    float* qs = big;
    float* ks = qs + J * LDQ;
    float* vs = ks + J * LDQ;
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.q_T, E, E, n, xs, E, acc);
        float b = __ldg(w.q_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) { qs[j * LDQ + n] = acc[j] + b; acc[j] = 0.f; }
        lin_tok<J>(w.k_T, E, E, n, xs, E, acc);
        b = __ldg(w.k_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) { ks[j * LDQ + n] = acc[j] + b; acc[j] = 0.f; }
        lin_tok<J>(w.v_T, E, E, n, xs, E, acc);
        b = __ldg(w.v_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) vs[j * LDQ + n] = acc[j] + b;
    }
    __syncthreads();
    {
        const float scale = (float)(1.0 / sqrt((double)HD));   // head_dims ** -0.5 (transformer.py:66)
        for (int t = tid; t < NH * J * J; t += E) {
            const int h = t / (J * J), r = t - h * J * J;
            const int jq = r / J, jk = r - jq * J;
            const float* qp = qs + jq * LDQ + h * HD;
            const float* kp = ks + jk * LDQ + h * HD;
            float dot = 0.f;
#pragma unroll 8
            for (int dd = 0; dd < HD; ++dd) dot = fmaf(qp[dd], kp[dd], dot);
            probs[(h * J + jq) * 16 + jk] = dot * scale;
        }
    }
    __syncthreads();
    for (int g = tid; g < NH * J; g += E) softmax_inplace(probs + g * 16, J);
    __syncthreads();
    {
        const int h = n / HD;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            float acc = 0.f;
            const float* pr = probs + (h * J + j) * 16;
#pragma unroll
            for (int jk = 0; jk < J; ++jk) acc = fmaf(pr[jk], vs[jk * LDQ + n], acc);
            ys[j * E + n] = acc;
        }
    }
    __syncthreads();
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.o_T, E, E, n, ys, E, acc);
        const float b = __ldg(w.o_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) big[j * E + n] = xs[j * E + n] + acc[j] + b;
    }
    __syncthreads();
    layernorm_rows<E>(big, xs, E, E, J, w.lns_w, w.lns_b);
    __syncthreads();

    // ---- A3: FFN ----
    float* hid = big;   // [J][FF]
    for (int m = tid; m < FF; m += E) {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.f1_T, E, FF, m, xs, E, acc);
        const float b = __ldg(w.f1_b + m);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float x = acc[j] + b;
            hid[j * FF + m] = 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
        }
    }
    __syncthreads();
    {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        lin_tok<J>(w.f2_T, FF, E, n, hid, FF, acc);
        const float b = __ldg(w.f2_b + n);
#pragma unroll
        for (int j = 0; j < J; ++j) ys[j * E + n] = xs[j * E + n] + acc[j] + b;
    }
    __syncthreads();
    layernorm_rows<E>(ys, xs, E, E, J, w.lnf_w, w.lnf_b);
    __syncthreads();
}

}  // namespace egr
