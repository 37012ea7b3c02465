// Batched token path: the small kernels between the token GEMMs (see token_batched.cuh).
#include "token_batched.cuh"

namespace egr {

// =====================================================================================================
// deformable sampling: one warp per (group, frame, joint, view, head)
//
// Lanes 0-15 each own one of the head's 16 points: softmax of the logits, the 4 bilinear corners (flat index, attention
// x bilinear weight, 0 for out-of-map corners) -> 64 (index, weight) pairs in shared memory.  The gather is branch-free
// (indices are clamped in-map, a zero weight contributes exactly 0) and split over the half-warps: each half takes every
// other corner, a lane loads 16 B = 8 of the 128 raw channels (and 8 B of the head's 64 position-table channels), so a
// warp-level load covers two corners and 16 corner loads are in flight per lane.  Halves are summed with one shuffle.
// =====================================================================================================
// volatile: the batch of gathers is issued back to back, ahead of the arithmetic that consumes it
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// two 16-bit values of the staged map -> fp32 (bf16: shifts; fp16: one cvt)
// fp16 pair of an fp32 value: hi = fp16(x), lo = fp16(x - hi) (operand rows [x_hi | x_lo] of the fp16 three-term product)
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
template <bool F16> __device__ __forceinline__ void cvt16x2(uint32_t w, float& lo, float& hi) {
    if (F16) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w)); lo = f.x; hi = f.y; }
    else { lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u); }
}

template <bool HAS_PTAB, int E, bool F16>
__global__ void __launch_bounds__(256, 4)
tok_sample_kernel(TokSampleArgs a) {
    constexpr int NH = TOK_NH, P = TOK_P, HD = E / NH, RAWC = TOK_RAWC;
    constexpr int EX = HAS_PTAB ? E : NH;
    static_assert(P == 16 && RAWC == 128 && (!HAS_PTAB || HD == 64), "lane mapping");
    __shared__ uint2 s_pairs[8][P * 4];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t unit = (int64_t)blockIdx.x * 8 + wib;
    const int64_t total = (int64_t)a.G * a.B * a.J * a.V * NH;
    if (unit >= total) return;
    // unit order (fastest first): head, refiner group, joint, view, frame.  The groups share the anchors, so the warps of
    // one CTA gather around the same point of the same view map, and one (frame, view) map (1 MB) is sampled by
    // neighbouring CTAs while it is resident in L2 / L1
    int64_t r = unit;
    const int h = (int)(r % NH); r /= NH;
    const int g = (int)(r % a.G); r /= a.G;
    const int j = (int)(r % a.J); r /= a.J;
    const int v = (int)(r % a.V);
    const int b = (int)(r / a.V);
    const int64_t t = (int64_t)b * a.J + j, T_ = (int64_t)a.B * a.J;
    const int LO = a.V * a.KA;                                   // offset of the x_lo half of a split row
    float* Arow = a.A + ((int64_t)g * T_ + t) * (int64_t)LO * (1 + a.split) + (int64_t)v * a.KA;
    const bool ok = a.valid[((int64_t)b * a.V + v) * a.J + j] != 0;
    // split == 2: the row is an fp16 pair [x_hi | x_lo] of 2 * LO halfs (operand of the fp16 three-term fold GEMM)
    __half* Hrow = reinterpret_cast<__half*>(a.A) + ((int64_t)g * T_ + t) * (int64_t)LO * 2 + (int64_t)v * a.KA;
    if (a.split == 2) {
        const __half zero = __float2half_rn(0.f);
        if (h == 0) {
            if (lane == 0) { Hrow[NH * RAWC + EX] = __float2half_rn(ok ? 1.f : 0.f); Hrow[LO + NH * RAWC + EX] = zero; }
            for (int c = NH * RAWC + EX + 1 + lane; c < a.KA; c += 32) { Hrow[c] = zero; Hrow[LO + c] = zero; }
        }
        if (!ok) {
            for (int s2 = 0; s2 < 2; ++s2) {
                __half* R = Hrow + s2 * LO;
                *reinterpret_cast<uint2*>(R + h * RAWC + lane * 4) = make_uint2(0u, 0u);
                if (HAS_PTAB) *reinterpret_cast<uint32_t*>(R + NH * RAWC + h * HD + lane * 2) = 0u;
                else if (lane == 0) R[NH * RAWC + h] = zero;
            }
            return;
        }
    } else
    if (h == 0) {     // validity column (carries output_proj's bias through the fold) + zero padding
        if (lane == 0) Arow[NH * RAWC + EX] = ok ? 1.f : 0.f;
        for (int c = NH * RAWC + EX + 1 + lane; c < a.KA; c += 32) Arow[c] = 0.f;
        if (a.split) for (int c = NH * RAWC + EX + lane; c < a.KA; c += 32) Arow[LO + c] = 0.f;
    }
    if (!ok && a.split != 2) {        // masked_fill(~anchors_valid, 0) after output_proj (:910 / :563): the whole row is zero
        for (int s2 = 0; s2 <= a.split; ++s2) {
            float* R = Arow + s2 * LO;
            *reinterpret_cast<float4*>(R + h * RAWC + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_PTAB) *reinterpret_cast<float2*>(R + NH * RAWC + h * HD + lane * 2) = make_float2(0.f, 0.f);
            else if (lane == 0) R[NH * RAWC + h] = 0.f;
        }
        return;
    }
    const float* oa = a.oa + ((int64_t)g * T_ + t) * TOK_OA;
    const int p = lane & (P - 1);
    // softmax over the 16 points of this head (deform_attn.py:125-130)
    const float logit = (lane < P) ? oa[NH * P * 2 + h * P + p] : -INFINITY;
    const float m = warp_max(logit);
    const float e = (lane < P) ? expf(logit - m) : 0.f;
    const float inv = 1.f / warp_sum(e);
    const float aw = e * inv;
    if (lane < P) {
        // loc = anchor + offset / (W, H)   (deform_attn.py:133-139), corners as in mmcv's kernel
        const float ax = a.anchors[(((int64_t)b * a.V + v) * a.J + j) * 2 + 0];
        const float ay = a.anchors[(((int64_t)b * a.V + v) * a.J + j) * 2 + 1];
        const float lx = ax + oa[h * P * 2 + p * 2 + 0] / (float)a.W;
        const float ly = ay + oa[h * P * 2 + p * 2 + 1] / (float)a.H;
        const Corners c = msda_corners(lx, ly, a.H, a.W);
#pragma unroll
        for (int q = 0; q < 4; ++q) s_pairs[wib][p * 4 + q] = make_uint2((uint32_t)c.idx[q], __float_as_uint(aw * c.w[q]));
    }
    __syncwarp();

    const int half = lane >> 4, l16 = lane & 15;
    const int64_t HW = (int64_t)a.H * a.W;
    const __nv_bfloat16* Xb = reinterpret_cast<const __nv_bfloat16*>(a.X) + ((int64_t)v * a.B + b) * HW * RAWC + l16 * 8;
    const __nv_bfloat16* pt = HAS_PTAB ? (a.ptab[g] + (int64_t)v * HW * E + h * HD + l16 * 4) : nullptr;
    float s8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float e4[4] = {0.f, 0.f, 0.f, 0.f};
    float wsum = 0.f;
    constexpr int UN = 8;                 // corner loads in flight per lane
#pragma unroll 1
    for (int i0 = 0; i0 < P * 2; i0 += UN) {
        uint4 x[UN];
        uint2 pu[UN];
        float cf[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const uint2 pr = s_pairs[wib][2 * (i0 + u) + half];
            cf[u] = __uint_as_float(pr.y);
            x[u] = ldg_nc_v4(Xb + (int64_t)pr.x * RAWC);
            if (HAS_PTAB) pu[u] = ldg_nc_v2(pt + (int64_t)pr.x * E);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const float coef = cf[u];
            float f0, f1;
            cvt16x2<F16>(x[u].x, f0, f1); s8[0] = fmaf(coef, f0, s8[0]); s8[1] = fmaf(coef, f1, s8[1]);
            cvt16x2<F16>(x[u].y, f0, f1); s8[2] = fmaf(coef, f0, s8[2]); s8[3] = fmaf(coef, f1, s8[3]);
            cvt16x2<F16>(x[u].z, f0, f1); s8[4] = fmaf(coef, f0, s8[4]); s8[5] = fmaf(coef, f1, s8[5]);
            cvt16x2<F16>(x[u].w, f0, f1); s8[6] = fmaf(coef, f0, s8[6]); s8[7] = fmaf(coef, f1, s8[7]);
            if (HAS_PTAB) {
                cvt16x2<F16>(pu[u].x, f0, f1); e4[0] = fmaf(coef, f0, e4[0]); e4[1] = fmaf(coef, f1, e4[1]);
                cvt16x2<F16>(pu[u].y, f0, f1); e4[2] = fmaf(coef, f0, e4[2]); e4[3] = fmaf(coef, f1, e4[3]);
            } else {
                wsum += coef;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s8[i] += __shfl_xor_sync(0xffffffffu, s8[i], 16);
    if (HAS_PTAB) {
#pragma unroll
        for (int i = 0; i < 4; ++i) e4[i] += __shfl_xor_sync(0xffffffffu, e4[i], 16);
    } else {
        wsum += __shfl_xor_sync(0xffffffffu, wsum, 16);
    }
    if (half == 0 && a.split == 2) {
        uint4 hi, lo;
        split_h2(s8[0], s8[1], hi.x, lo.x); split_h2(s8[2], s8[3], hi.y, lo.y);
        split_h2(s8[4], s8[5], hi.z, lo.z); split_h2(s8[6], s8[7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(Hrow + h * RAWC + l16 * 8) = hi;
        *reinterpret_cast<uint4*>(Hrow + LO + h * RAWC + l16 * 8) = lo;
        if (HAS_PTAB) {
            uint2 eh, el;
            split_h2(e4[0], e4[1], eh.x, el.x); split_h2(e4[2], e4[3], eh.y, el.y);
            *reinterpret_cast<uint2*>(Hrow + NH * RAWC + h * HD + l16 * 4) = eh;
            *reinterpret_cast<uint2*>(Hrow + LO + NH * RAWC + h * HD + l16 * 4) = el;
        } else if (l16 == 0) {
            const __half wh = __float2half_rn(wsum);
            Hrow[NH * RAWC + h] = wh;
            Hrow[LO + NH * RAWC + h] = __float2half_rn(wsum - __half2float(wh));
        }
    } else if (half == 0) {
        float4* o = reinterpret_cast<float4*>(Arow + h * RAWC + l16 * 8);
        if (a.split) {
            o[0] = make_float4(s8[0], s8[1], s8[2], s8[3]);
            o[1] = make_float4(s8[4], s8[5], s8[6], s8[7]);
            float4* ol = reinterpret_cast<float4*>(Arow + LO + h * RAWC + l16 * 8);
            ol[0] = make_float4(tf32_lo(s8[0]), tf32_lo(s8[1]), tf32_lo(s8[2]), tf32_lo(s8[3]));
            ol[1] = make_float4(tf32_lo(s8[4]), tf32_lo(s8[5]), tf32_lo(s8[6]), tf32_lo(s8[7]));
            if (HAS_PTAB) {
                *reinterpret_cast<float4*>(Arow + NH * RAWC + h * HD + l16 * 4) = make_float4(e4[0], e4[1], e4[2], e4[3]);
                *reinterpret_cast<float4*>(Arow + LO + NH * RAWC + h * HD + l16 * 4) =
                    make_float4(tf32_lo(e4[0]), tf32_lo(e4[1]), tf32_lo(e4[2]), tf32_lo(e4[3]));
            } else if (l16 == 0) { Arow[NH * RAWC + h] = wsum; Arow[LO + NH * RAWC + h] = tf32_lo(wsum); }
        } else {
            o[0] = make_float4(round_tf32(s8[0]), round_tf32(s8[1]), round_tf32(s8[2]), round_tf32(s8[3]));
            o[1] = make_float4(round_tf32(s8[4]), round_tf32(s8[5]), round_tf32(s8[6]), round_tf32(s8[7]));
            if (HAS_PTAB)
                *reinterpret_cast<float4*>(Arow + NH * RAWC + h * HD + l16 * 4) =
                    make_float4(round_tf32(e4[0]), round_tf32(e4[1]), round_tf32(e4[2]), round_tf32(e4[3]));
            else if (l16 == 0) Arow[NH * RAWC + h] = round_tf32(wsum);
        }
    }
}

int tok_sample(const TokSampleArgs& a, int act_bf16, cudaStream_t st) {
    const bool ptab = a.ptab != nullptr;
    EGR_CHECK((a.E == 256 && ptab) || (a.E == 128 && !ptab), EGR_ERR_UNSUPPORTED, "tok_sample: E=%d ptab=%d", a.E, (int)ptab);
    EGR_CHECK(a.KA == tok_ka(a.E, ptab), EGR_ERR_INVALID, "tok_sample: KA=%d", a.KA);
    EGR_CHECK(act_bf16 == 1 || act_bf16 == 2, EGR_ERR_UNSUPPORTED, "tok_sample: the batched token path samples a 16-bit channels-last copy");
    const int64_t total = (int64_t)a.G * a.B * a.J * a.V * TOK_NH;
    const int blocks = (int)ceil_div64(total, 8);
    if (act_bf16 == 2) {
        if (ptab) EGR_LAUNCH((tok_sample_kernel<true, 256, true>), blocks, 256, 0, st, a);
        else EGR_LAUNCH((tok_sample_kernel<false, 128, true>), blocks, 256, 0, st, a);
    } else {
        if (ptab) EGR_LAUNCH((tok_sample_kernel<true, 256, false>), blocks, 256, 0, st, a);
        else EGR_LAUNCH((tok_sample_kernel<false, 128, false>), blocks, 256, 0, st, a);
    }
    return EGR_OK;
}

// =====================================================================================================
// joint self-attention: one CTA per frame, one warp per head
// =====================================================================================================
template <int E>
__global__ void __launch_bounds__(128)
tok_attn_kernel(const float* __restrict__ qkv, float* __restrict__ o, int J, int split) {
    constexpr int NH = TOK_NH, HD = E / NH, LD = HD + 1, MJ = 16;
    extern __shared__ float sm[];
    pdl_trigger();
    pdl_wait();
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* q = sm + h * (3 * MJ * LD + MJ * MJ);
    float* k = q + MJ * LD;
    float* v = k + MJ * LD;
    float* pr = v + MJ * LD;      // [MJ][MJ]
    const int64_t t0 = (int64_t)blockIdx.x * J;
    for (int i = lane; i < J * HD; i += 32) {
        const int j = i / HD, d = i - j * HD;
        const float* row = qkv + (t0 + j) * 3 * E + h * HD + d;
        q[j * LD + d] = row[0];
        k[j * LD + d] = row[E];
        v[j * LD + d] = row[2 * E];
    }
    __syncwarp();
    const float scale = (float)(1.0 / sqrt((double)HD));      // head_dims ** -0.5 (transformer.py:66)
    const int jq = lane & 15, kh = lane >> 4;
    float sc[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int jk = kh * 8 + i;
        float dot = 0.f;
        if (jq < J && jk < J) {
#pragma unroll 8
            for (int d = 0; d < HD; ++d) dot = fmaf(q[jq * LD + d], k[jk * LD + d], dot);
            sc[i] = dot * scale;
            mx = fmaxf(mx, sc[i]);
        } else {
            sc[i] = -INFINITY;
        }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = (sc[i] == -INFINITY) ? 0.f : expf(sc[i] - mx); s += sc[i]; }
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    const float inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < 8; ++i) pr[jq * MJ + kh * 8 + i] = sc[i] * inv;
    __syncwarp();
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int d = lane; d < HD; d += 32) {
            float acc = 0.f;
            for (int jk = 0; jk < J; ++jk) acc = fmaf(pr[j * MJ + jk], v[jk * LD + d], acc);
            tok_operand_store(o + (t0 + j) * E * (1 + split), h * HD + d, E, acc, split);
        }
    }
}

int tok_attn(const float* qkv, float* o, int n_frames, int J, int E, cudaStream_t st, int split) {
    EGR_CHECK(J <= 16 && (E == 128 || E == 256), EGR_ERR_UNSUPPORTED, "tok_attn: J=%d E=%d", J, E);
    const int HD = E / TOK_NH;
    const size_t smem = sizeof(float) * TOK_NH * (3 * 16 * (HD + 1) + 16 * 16);
    if (E == 256) {
        auto kfn = tok_attn_kernel<256>;
        EGR_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EGR_LAUNCH(kfn, n_frames, 128, smem, st, qkv, o, J, split);
    } else {
        auto kfn = tok_attn_kernel<128>;
        EGR_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EGR_LAUNCH(kfn, n_frames, 128, smem, st, qkv, o, J, split);
    }
    return EGR_OK;
}

// =====================================================================================================
// residual + LayerNorm (eps 1e-5, biased variance), one warp per token row
// =====================================================================================================
template <int E>
__device__ __forceinline__ void ln_row(const float* v_in, float (&v)[E / 32], float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < E / 32; ++i) s += v[i];
    mean = warp_sum(s) * (1.f / E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < E / 32; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    rstd = rsqrtf(warp_sum(q) * (1.f / E) + 1e-5f);
}

template <int E>
__global__ void __launch_bounds__(256)
tok_add_ln_kernel(const float* __restrict__ res, const float* __restrict__ z, float* __restrict__ out, int64_t rows,
                  int rows_per_group, const float* const* __restrict__ gamma, const float* const* __restrict__ beta, int split,
                  float* __restrict__ out2, const float* const* __restrict__ gamma2, const float* const* __restrict__ beta2) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t ldr = (int64_t)E * (1 + split);      // res / out rows: [x | x_lo] when split
    if (row >= rows) return;
    const int g = (int)(row / rows_per_group);
    const float* ga = gamma[g];
    const float* be = beta[g];
    float v[E / 32];
#pragma unroll
    for (int i = 0; i < E / 32; ++i) {
        const int c = lane + 32 * i;
        v[i] = z[row * E + c] + (res ? res[row * ldr + c] : 0.f);
    }
    float mean, rstd;
    ln_row<E>(nullptr, v, mean, rstd);
#pragma unroll
    for (int i = 0; i < E / 32; ++i) {
        const int c = lane + 32 * i;
        v[i] = (v[i] - mean) * rstd * __ldg(ga + c) + __ldg(be + c);
        tok_operand_store(out + row * ldr, c, E, v[i], split);
    }
    if (out2) {      // a second LayerNorm of the row just written (pose3d: post_norm[l] of the layer output), as stored
        const float* g2 = gamma2[g];
        const float* b2 = beta2[g];
#pragma unroll
        for (int i = 0; i < E / 32; ++i) v[i] = split ? v[i] : round_tf32(v[i]);
        ln_row<E>(nullptr, v, mean, rstd);
#pragma unroll
        for (int i = 0; i < E / 32; ++i) {
            const int c = lane + 32 * i;
            tok_operand_store(out2 + row * ldr, c, E, (v[i] - mean) * rstd * __ldg(g2 + c) + __ldg(b2 + c), split);
        }
    }
}

int tok_add_ln(const float* res, const float* z, float* out, int G, int rows_per_group, int E, const float* const* gamma,
               const float* const* beta, cudaStream_t st, int split, float* out2, const float* const* gamma2,
               const float* const* beta2) {
    const int64_t rows = (int64_t)G * rows_per_group;
    const int blocks = (int)ceil_div64(rows, 8);
    if (E == 256) EGR_LAUNCH(tok_add_ln_kernel<256>, blocks, 256, 0, st, res, z, out, rows, rows_per_group, gamma, beta, split, out2, gamma2, beta2);
    else if (E == 128) EGR_LAUNCH(tok_add_ln_kernel<128>, blocks, 256, 0, st, res, z, out, rows, rows_per_group, gamma, beta, split, out2, gamma2, beta2);
    else return fail(EGR_ERR_UNSUPPORTED, "tok_add_ln: E=%d", E);
    return EGR_OK;
}

// post_norm + [J][16x16] token image written transposed: xT[pos][j], j padded to 16 (egoposeformer_heatmap_mvf_ex.py:707-711)
template <typename T>
__global__ void __launch_bounds__(256)
tok_ln_image_kernel(const float* __restrict__ x, T* __restrict__ xT, int B, int J, const float* const* __restrict__ gamma,
                    const float* const* __restrict__ beta) {
    constexpr int E = 256;
    __shared__ float ys[16 * E];
    pdl_trigger();
    pdl_wait();
    const int g = blockIdx.x / B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* xin = x + (int64_t)blockIdx.x * J * E;
    const float* ga = gamma[g];
    const float* be = beta[g];
    for (int j = warp; j < J; j += 8) {
        float v[E / 32];
#pragma unroll
        for (int i = 0; i < E / 32; ++i) v[i] = xin[j * E + lane + 32 * i];
        float mean, rstd;
        ln_row<E>(nullptr, v, mean, rstd);
#pragma unroll
        for (int i = 0; i < E / 32; ++i) {
            const int c = lane + 32 * i;
            ys[j * E + c] = (v[i] - mean) * rstd * __ldg(ga + c) + __ldg(be + c);
        }
    }
    __syncthreads();
    T* out = xT + (int64_t)blockIdx.x * E * 16;
    for (int i = threadIdx.x; i < E * 16; i += 256) {
        const int pos = i >> 4, j = i & 15;
        ActT<T>::st(out + i, (j < J) ? ys[j * E + pos] : 0.f);
    }
}

int tok_ln_image(const float* x, void* xT, int xT_bf16, int G, int B, int J, int E, const float* const* gamma,
                 const float* const* beta, cudaStream_t st) {
    EGR_CHECK(E == 256 && J <= 16, EGR_ERR_UNSUPPORTED, "tok_ln_image: E=%d J=%d", E, J);
    if (xT_bf16) EGR_LAUNCH(tok_ln_image_kernel<__nv_bfloat16>, G * B, 256, 0, st, x, (__nv_bfloat16*)xT, B, J, gamma, beta);
    else EGR_LAUNCH(tok_ln_image_kernel<float>, G * B, 256, 0, st, x, (float*)xT, B, J, gamma, beta);
    return EGR_OK;
}

// =====================================================================================================
// mvfex jqa query input (HeatmapMVF.forward :655-665): x0 = embed_j + fc_bfb(avgpool(bfb)) + heatmap_proj(heatmap_j)
// =====================================================================================================
__global__ void __launch_bounds__(256)
tok_avgpool_kernel(const float* __restrict__ bfb, int64_t bfb_bs, int64_t bfb_gs, int hw, float* __restrict__ pooled, int B, int C, int split) {
    // hw == 64: a half-warp per channel, one float4 per lane (256 B per channel, coalesced), 4 shuffle steps
    pdl_trigger();
    pdl_wait();
    const int g = blockIdx.x / B, b = blockIdx.x - g * B;
    const float* bf = bfb + (int64_t)g * bfb_gs + (int64_t)b * bfb_bs;
    const int l16 = threadIdx.x & 15, hwarp = threadIdx.x >> 4;
    const int c0 = blockIdx.y * (C / gridDim.y), c1 = c0 + C / gridDim.y;
    for (int c = c0 + hwarp; c < c1; c += 16) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(bf + (int64_t)c * hw) + l16);
        float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (l16 == 0) tok_operand_store(pooled + (int64_t)blockIdx.x * C * (1 + split), c, C, s / (float)hw, split);
    }
}
int tok_avgpool(const float* bfb, int64_t bfb_bs, int64_t bfb_gs, int hw, float* pooled, int G, int B, int C, cudaStream_t st, int split) {
    EGR_CHECK(hw == 64 && C % 64 == 0, EGR_ERR_UNSUPPORTED, "tok_avgpool: hw=%d C=%d (stride-32 map of a 256x256 image is 8x8)", hw, C);
    EGR_LAUNCH(tok_avgpool_kernel, dim3(G * B, 4), 256, 0, st, bfb, bfb_bs, bfb_gs, hw, pooled, B, C, split);
    return EGR_OK;
}

__global__ void __launch_bounds__(256)
tok_add_query_kernel(const float* __restrict__ y0, const float* __restrict__ vb, const float* const* __restrict__ jq,
                     float* __restrict__ x0, int B, int J, int E, int split) {
    pdl_trigger();
    pdl_wait();
    const int g = blockIdx.x / B;
    const float* q = jq[g];
    const int64_t base = (int64_t)blockIdx.x * J * E;
    for (int i = threadIdx.x; i < J * E; i += 256) {
        const int j = i / E, n = i - j * E;
        tok_operand_store(x0 + ((int64_t)blockIdx.x * J + j) * E * (1 + split), n, E,
                          __ldg(q + i) + vb[(int64_t)blockIdx.x * E + n] + y0[base + i], split);
    }
}
int tok_add_query(const float* y0, const float* vb, const float* const* jq, float* x0, int G, int B, int J, int E, cudaStream_t st, int split) {
    EGR_LAUNCH(tok_add_query_kernel, G * B, 256, 0, st, y0, vb, jq, x0, B, J, E, split);
    return EGR_OK;
}

__global__ void __launch_bounds__(256)
tok_make_lo_kernel(float* __restrict__ buf, int64_t n, int K) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 of the x half
    if (i * 4 >= n) return;
    const int64_t e = i * 4, r = e / K;
    const int c = (int)(e - r * K);
    const float4 x = *reinterpret_cast<const float4*>(buf + r * 2 * K + c);
    *reinterpret_cast<float4*>(buf + r * 2 * K + K + c) = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
}
int tok_make_lo(float* buf, int64_t rows, int K, cudaStream_t st) {
    EGR_CHECK(K % 4 == 0, EGR_ERR_UNSUPPORTED, "tok_make_lo: K=%d", K);
    const int64_t n = rows * K;
    EGR_LAUNCH(tok_make_lo_kernel, (int)ceil_div64(n / 4, 256), 256, 0, st, buf, n, K);
    return EGR_OK;
}

// =====================================================================================================
// pose3d: reprojection (P3) + query_gen_mlp.0 + ReLU           egoposeformer_mvf_ex.py:340-382, :400-406
// =====================================================================================================
__global__ void __launch_bounds__(128)
pose_query0_kernel(PoseQueryArgs a) {
    constexpr int E = 128, MJ = 16;
    __shared__ float s_p3[MJ * 4];
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x, tid = threadIdx.x, J = a.J;
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
            a.anchors[(((int64_t)b * a.V + v) * J + j) * 2 + 0] = u;
            a.anchors[(((int64_t)b * a.V + v) * J + j) * 2 + 1] = vv;
            a.valid[((int64_t)b * a.V + v) * J + j] = fov ? 1 : 0;
        }
        s_p3[j * 4 + 0] = (float)(j + 1) / (float)J;
        s_p3[j * 4 + 1] = px; s_p3[j * 4 + 2] = py; s_p3[j * 4 + 3] = pz;
        float* o = a.p3 + ((int64_t)b * J + j) * 4;
        o[0] = s_p3[j * 4 + 0]; o[1] = px; o[2] = py; o[3] = pz;
    }
    __syncthreads();
    const int n = tid;
    const float w0 = __ldg(a.g0_T + 0 * E + n), w1 = __ldg(a.g0_T + 1 * E + n);
    const float w2 = __ldg(a.g0_T + 2 * E + n), w3 = __ldg(a.g0_T + 3 * E + n);
    const float b0 = __ldg(a.g0_b + n);
    for (int j = 0; j < J; ++j) {
        float acc = 0.f;
        acc = fmaf(w0, s_p3[j * 4 + 0], acc);
        acc = fmaf(w1, s_p3[j * 4 + 1], acc);
        acc = fmaf(w2, s_p3[j * 4 + 2], acc);
        acc = fmaf(w3, s_p3[j * 4 + 3], acc);
        a.x0[((int64_t)b * J + j) * E + n] = round_tf32(fmaxf(acc + b0, 0.f));
    }
}

int pose_query0(const PoseQueryArgs& a, cudaStream_t st) {
    EGR_CHECK(a.E == 128 && a.J <= 16 && a.V <= 4, EGR_ERR_UNSUPPORTED, "pose_query0: E=%d J=%d V=%d", a.E, a.J, a.V);
    EGR_LAUNCH(pose_query0_kernel, a.B, 128, 0, st, a);
    return EGR_OK;
}

__global__ void pose_reg_out_kernel(const float* __restrict__ r, const float* __restrict__ w2_T, const float* __restrict__ b2,
                                    const float* __restrict__ p3, float* __restrict__ preds, int T, int E) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * 3) return;
    const int t = i / 3, c = i - t * 3;
    float acc = __ldg(b2 + c);
    for (int k = 0; k < E; ++k) acc = fmaf(__ldg(w2_T + k * 3 + c), r[(int64_t)t * E + k], acc);
    preds[i] = acc + p3[(int64_t)t * 4 + 1 + c];
}

int pose_reg_out(const float* r, const float* w2_T, const float* b2, const float* p3, float* preds, int T, int E, cudaStream_t st) {
    EGR_LAUNCH(pose_reg_out_kernel, ceil_div(T * 3, 128), 128, 0, st, r, w2_T, b2, p3, preds, T, E);
    return EGR_OK;
}

// =====================================================================================================
__global__ void small_matmul_ex_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                                       const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N;
        const int n = (int)(i - m * N);
        float acc = bias ? bias[n] : 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(A[m * lda + k], B[(int64_t)k * ldb + n], acc);
        C[m * ldc + n] = acc;
    }
}
int small_matmul_ex(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                    int M, int N, int K, cudaStream_t st) {
    const int64_t g = ceil_div64((int64_t)M * N, 256);
    small_matmul_ex_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K);
    EGR_LAUNCHED();
    return EGR_OK;
}

}  // namespace egr
