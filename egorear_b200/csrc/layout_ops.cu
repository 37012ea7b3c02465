// HBM-bound layout / resampling stages: NCHW->NHWC staging, bilinear x2 (align_corners=True) tails of the
// heatmap heads and of the refined-feature projection, 2x2 max-pool, and one-time weight repacking.
#include "layout_ops.cuh"
#include <type_traits>

namespace egr {
int g_opt_tail_mma = 1;     // option "tail_mma": head tails on head_tail_mma.cu (0 = head_tail_tc.cu)


extern int g_opt_tc;

struct UpC {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ UpC upc(int dst, int in_size) {
    // nn.Upsample(scale_factor=2, bilinear, align_corners=True): src = dst * (in-1)/(out-1), out = 2*in
    const float scale = (float)(in_size - 1) / (float)(2 * in_size - 1);
    const float s = scale * (float)dst;
    UpC c;
    c.i0 = (int)s;
    c.i1 = c.i0 + ((c.i0 < in_size - 1) ? 1 : 0);
    c.l1 = s - (float)c.i0;
    c.l0 = 1.f - c.l1;
    return c;
}

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> view-major NHWC (float | bf16)
// ---------------------------------------------------------------------------------------------
template <typename T, bool RND>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int B, int V, int C, int HW) {
    extern __shared__ float tile[];   // [C][33]
    pdl_trigger();
    pdl_wait();
    const int p0 = blockIdx.x * 32;
    const int bv = blockIdx.y;        // b * V + v
    const int b = bv / V, v = bv - b * V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* src = in + (int64_t)bv * C * HW + p0;
    for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = __ldcs(src + (int64_t)c * HW + lane);
    __syncthreads();
    T* dst = out + (((int64_t)v * B + b) * HW + p0) * C;
    for (int p = warp; p < 32; p += 8)
        for (int c = lane; c < C; c += 32) {
            const float x = tile[c * 33 + p];
            ActT<T>::st(dst + (int64_t)p * C + c, RND ? round_tf32(x) : x);
        }
}

// 16-bit output (bf16 | fp16), C = 128: 64-pixel tiles.  Two channel rows are packed into 16x2 words on the way into shared memory
// (tile[pair][pixel], odd stride), so the write phase reads conflict-free and every store instruction writes one full
// 128 B line of a pixel's 256 B channel vector; reads are 2 x 128 B per channel row.
constexpr int NT_PX = 64, NT_C = 128, NT_LD = NT_PX + 1;
template <typename T16> __device__ __forceinline__ uint32_t pack16x2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack16x2<__nv_bfloat16>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <> __device__ __forceinline__ uint32_t pack16x2<__half>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <typename T16>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_16_kernel(const float* __restrict__ in, T16* __restrict__ out, int B, int V, int HW) {
    __shared__ uint32_t tile[(NT_C / 2) * NT_LD];
    pdl_trigger();
    pdl_wait();
    const int p0 = blockIdx.x * NT_PX;
    const int bv = blockIdx.y;        // b * V + v
    const int b = bv / V, v = bv - b * V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* src = in + (int64_t)bv * NT_C * HW + p0 + lane;
#pragma unroll 4
    for (int cp = warp; cp < NT_C / 2; cp += 8) {
        const float* r0 = src + (int64_t)(2 * cp) * HW;
        const float a0 = __ldcs(r0), a1 = __ldcs(r0 + 32), b0 = __ldcs(r0 + HW), b1 = __ldcs(r0 + HW + 32);
        tile[cp * NT_LD + lane] = pack16x2<T16>(a0, b0);
        tile[cp * NT_LD + lane + 32] = pack16x2<T16>(a1, b1);
    }
    __syncthreads();
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (((int64_t)v * B + b) * HW + p0) * NT_C);
#pragma unroll 4
    for (int p = warp; p < NT_PX; p += 8) {
        dst[(int64_t)p * (NT_C / 2) + lane] = tile[lane * NT_LD + p];
        dst[(int64_t)p * (NT_C / 2) + lane + 32] = tile[(lane + 32) * NT_LD + p];
    }
}

int nchw_to_nhwc(const float* in, void* out, int B, int V, int C, int HW, int out_mode, cudaStream_t st) {
    if ((out_mode == 1 || out_mode == 3) && C == NT_C && HW % NT_PX == 0) {
        if (out_mode == 1) EGR_LAUNCH(nchw_to_nhwc_16_kernel<__nv_bfloat16>, dim3(HW / NT_PX, B * V), 256, 0, st, in, (__nv_bfloat16*)out, B, V, HW);
        else EGR_LAUNCH(nchw_to_nhwc_16_kernel<__half>, dim3(HW / NT_PX, B * V), 256, 0, st, in, (__half*)out, B, V, HW);
        return EGR_OK;
    }
    EGR_CHECK(HW % 32 == 0 && C * 33 * 4 <= 48 * 1024, EGR_ERR_UNSUPPORTED, "nchw_to_nhwc: HW=%d C=%d", HW, C);
    dim3 grid(HW / 32, B * V);
    const size_t smem = sizeof(float) * C * 33;
    if (out_mode == 1) EGR_LAUNCH((nchw_to_nhwc_kernel<__nv_bfloat16, false>), grid, 256, smem, st, in, (__nv_bfloat16*)out, B, V, C, HW);
    else if (out_mode == 3) EGR_LAUNCH((nchw_to_nhwc_kernel<__half, false>), grid, 256, smem, st, in, (__half*)out, B, V, C, HW);
    else if (out_mode == 2) EGR_LAUNCH((nchw_to_nhwc_kernel<float, true>), grid, 256, smem, st, in, (float*)out, B, V, C, HW);
    else EGR_LAUNCH((nchw_to_nhwc_kernel<float, false>), grid, 256, smem, st, in, (float*)out, B, V, C, HW);
    return EGR_OK;
}

// ---------------------------------------------------------------------------------------------
// shared tile loader: source rows [sr0, sr0+nsr) of z[img] ([Hs*Ws][C]) -> smem s[c][nsr*Ws] (stride ld)
// ---------------------------------------------------------------------------------------------
constexpr int STRIP = 4;   // output rows per block

template <typename TZ>
__device__ __forceinline__ void load_src_rows(const TZ* __restrict__ zimg, float* s, int ld, int sr0, int nsr, int Ws,
                                              int C) {
    const int npix = nsr * Ws;
    const TZ* src = zimg + (int64_t)sr0 * Ws * C;
    for (int i = threadIdx.x; i < npix * C; i += blockDim.x) {
        const int pix = i / C, c = i - pix * C;
        s[c * ld + pix] = ActT<TZ>::ld(src + i);
    }
}

__device__ __forceinline__ float interp(const float* sc, int Ws, int sr0, const UpC& cy, const UpC& cx) {
    const float* r0 = sc + (cy.i0 - sr0) * Ws;
    const float* r1 = sc + (cy.i1 - sr0) * Ws;
    return cy.l0 * (cx.l0 * r0[cx.i0] + cx.l1 * r0[cx.i1]) + cy.l1 * (cx.l0 * r1[cx.i0] + cx.l1 * r1[cx.i1]);
}

// ---------------------------------------------------------------------------------------------
// head tail: up2 + ReLU + 1x1 C->J
// ---------------------------------------------------------------------------------------------
constexpr int HJ = 16;

template <typename TZ, typename TH>
__global__ void __launch_bounds__(256)
head_up_conv_kernel(const TZ* __restrict__ z, const float* __restrict__ w, const float* __restrict__ bias, int4 wsel,
                    int B, int Hs, int Ws, int C, int J, float* __restrict__ hm, int64_t hm_bs, int64_t hm_gs,
                    TH* __restrict__ hm_t) {
    extern __shared__ float smem[];
    const int ld = STRIP * Ws + 1;
    float* s = smem;                 // [C][ld]
    float* sw = smem + C * ld;       // [C][HJ]
    __shared__ float sb[HJ];

    const int y0 = blockIdx.x * STRIP;
    const int g = blockIdx.y / B, b = blockIdx.y - g * B;
    const int img = g * B + b;
    const int sel = (g == 0) ? wsel.x : (g == 1) ? wsel.y : (g == 2) ? wsel.z : wsel.w;
    const float* wv = w + (int64_t)sel * J * C;
    for (int i = threadIdx.x; i < C * HJ; i += blockDim.x) {
        const int c = i / HJ, j = i - c * HJ;
        sw[i] = (j < J) ? wv[j * C + c] : 0.f;
    }
    if (threadIdx.x < HJ) sb[threadIdx.x] = (threadIdx.x < J) ? bias[(int64_t)sel * J + threadIdx.x] : 0.f;

    const int H = 2 * Hs, W = 2 * Ws;
    const int sr0 = upc(y0, Hs).i0;
    const int sr1 = upc(min(y0 + STRIP - 1, H - 1), Hs).i1;
    load_src_rows<TZ>(z + (int64_t)img * Hs * Ws * C, s, ld, sr0, sr1 - sr0 + 1, Ws, C);
    __syncthreads();

    for (int px = threadIdx.x; px < STRIP * W; px += blockDim.x) {
        const int yy = px / W, x = px - yy * W;
        const int y = y0 + yy;
        if (y >= H) break;
        const UpC cy = upc(y, Hs), cx = upc(x, Ws);
        float acc[HJ];
#pragma unroll
        for (int j = 0; j < HJ; ++j) acc[j] = sb[j];
        for (int c = 0; c < C; ++c) {
            const float u = fmaxf(interp(s + c * ld, Ws, sr0, cy, cx), 0.f);
            const float4* wr = reinterpret_cast<const float4*>(sw + c * HJ);
#pragma unroll
            for (int q = 0; q < HJ / 4; ++q) {
                const float4 ww = wr[q];
                acc[q * 4 + 0] = fmaf(ww.x, u, acc[q * 4 + 0]);
                acc[q * 4 + 1] = fmaf(ww.y, u, acc[q * 4 + 1]);
                acc[q * 4 + 2] = fmaf(ww.z, u, acc[q * 4 + 2]);
                acc[q * 4 + 3] = fmaf(ww.w, u, acc[q * 4 + 3]);
            }
        }
        float* o = hm + (int64_t)b * hm_bs + (int64_t)g * hm_gs + (int64_t)y * W + x;
#pragma unroll
        for (int j = 0; j < HJ; ++j)
            if (j < J) o[(int64_t)j * H * W] = acc[j];
        if (hm_t) {
            TH* ot = hm_t + ((int64_t)img * J) * H * W + (int64_t)y * W + x;
#pragma unroll
            for (int j = 0; j < HJ; ++j)
                if (j < J) ActT<TH>::st(ot + (int64_t)j * H * W, acc[j]);
        }
    }
}

// =============================================================================================
// Fast paths for the shipped geometry: 32x32 -> 64x64, C = 128 (every use on the hot path).
// One block = one image x 4 output rows; the (at most) 4 source rows they touch live in shared memory.
// Bilinear coefficients come from a 64-entry table built with the same fp32 formula as upc().
// =============================================================================================
constexpr int FS = 32, FO = 64, FCH = 128, FSTRIP = 4, FROWS = 4;

struct UpTab {
    int8_t i0[FO], i1[FO];
    float l0[FO], l1[FO];
};
__device__ __forceinline__ void build_uptab(UpTab* t) {
    if (threadIdx.x < FO) {
        const UpC c = upc(threadIdx.x, FS);
        t->i0[threadIdx.x] = (int8_t)c.i0; t->i1[threadIdx.x] = (int8_t)c.i1;
        t->l0[threadIdx.x] = c.l0; t->l1[threadIdx.x] = c.l1;
    }
}
template <typename T> __device__ __forceinline__ void unpack8(const T* p, float* v);
template <> __device__ __forceinline__ void unpack8<float>(const float* p, float* v) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <typename T> __device__ __forceinline__ void pack8_store(T* p, const float* v);
template <> __device__ __forceinline__ void pack8_store<float>(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void pack8_store<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
}

// R1 tail, fast: relu(up2(z)) -> fp32 NCHW (float4 streaming stores) and up to two channels-last copies (16-byte stores)
// in bf16, fp16 or fp32 rounded to TF32.  With fp16 z (the tensor-core path) the channels-last pass interpolates in half2
// arithmetic (HFMA2: two channels per instruction, no unpacking) - z already carries 16-bit rounding and the result is
// stored in 16 bits, so fp16 arithmetic loses nothing that matters; the kernel is then bound by its stores.
template <> __device__ __forceinline__ void unpack8<__half>(const __half* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void pack8_store<__half>(__half* p, const float* v) {
    uint4 u;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.y) : "f"(v[3]), "f"(v[2]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.z) : "f"(v[5]), "f"(v[4]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.w) : "f"(v[7]), "f"(v[6]));
    *reinterpret_cast<uint4*>(p) = u;
}
// store 8 channels-last values in the requested type: 0 fp32 rounded to TF32, 1 bf16, 2 fp16, 3 fp32 as is
__device__ __forceinline__ void store8_dt(void* base, int64_t off, int dt, float* r) {
    if (dt == 1) pack8_store<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(base) + off, r);
    else if (dt == 2) pack8_store<__half>(reinterpret_cast<__half*>(base) + off, r);
    else if (dt == 3) pack8_store<float>(reinterpret_cast<float*>(base) + off, r);
    else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = round_tf32(r[i]);
        pack8_store<float>(reinterpret_cast<float*>(base) + off, r);
    }
}

template <typename TZ>
__global__ void __launch_bounds__(256)
up2_relu_dual_fast_kernel(const TZ* __restrict__ z, int B, float* __restrict__ out_nchw, int64_t o_bs, int64_t o_gs,
                          void* __restrict__ cl0, int cl0_dt, void* __restrict__ cl1, int cl1_dt, int fp32_interp) {
    constexpr int SLD = 130;                               // transposed tile row stride: conflict-free both ways
    constexpr bool HALF = std::is_same<TZ, __half>::value;
    extern __shared__ __align__(16) uint8_t fsm[];
    TZ* s1 = reinterpret_cast<TZ*>(fsm);                   // [FROWS*FS px][FCH]   as loaded
    TZ* s2 = s1 + FROWS * FS * FCH;                        // [FCH][SLD]           transposed
    __shared__ UpTab tab;
    pdl_trigger();
    const int y0 = blockIdx.x * FSTRIP;
    const int g = blockIdx.y / B, b = blockIdx.y - g * B;
    const int img = g * B + b;
    const int sr0 = upc(y0, FS).i0;
    const int nsr = upc(y0 + FSTRIP - 1, FS).i1 - sr0 + 1;
    const int npix = nsr * FS;
    build_uptab(&tab);
    pdl_wait();
    {
        const uint4* src = reinterpret_cast<const uint4*>(z + ((int64_t)img * FS * FS + (int64_t)sr0 * FS) * FCH);
        uint4* dst = reinterpret_cast<uint4*>(s1);
        const int n16 = npix * FCH * (int)sizeof(TZ) / 16;
        for (int i = threadIdx.x; i < n16; i += 256) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (out_nchw) {       // the transposed copy only serves the NCHW pass
        for (int i = threadIdx.x; i < npix * FCH; i += 256) {
            const int p = i / FCH, c = i - p * FCH;
            s2[c * SLD + p] = s1[i];
        }
    }
    // ---- channels-last copies: item = (pixel, 8 channels) ----
    if (cl0 || cl1) {
        const int64_t base = ((int64_t)img * FO * FO + (int64_t)y0 * FO) * FCH;
#pragma unroll 2
        for (int it = threadIdx.x; it < FSTRIP * FO * (FCH / 8); it += 256) {
            const int c8 = it & 15, px = it >> 4;
            const int yy = px >> 6, x = px & 63;
            const int r0 = tab.i0[y0 + yy] - sr0, r1 = tab.i1[y0 + yy] - sr0;
            const float ly0 = tab.l0[y0 + yy], ly1 = tab.l1[y0 + yy];
            const int xa = tab.i0[x], xb = tab.i1[x];
            const float lx0 = tab.l0[x], lx1 = tab.l1[x];
            const int64_t off = base + (int64_t)px * FCH + c8 * 8;
            if (HALF && !fp32_interp) {
                const uint4 ua = *reinterpret_cast<const uint4*>(s1 + (r0 * FS + xa) * FCH + c8 * 8);
                const uint4 ub = *reinterpret_cast<const uint4*>(s1 + (r0 * FS + xb) * FCH + c8 * 8);
                const uint4 uc = *reinterpret_cast<const uint4*>(s1 + (r1 * FS + xa) * FCH + c8 * 8);
                const uint4 ud = *reinterpret_cast<const uint4*>(s1 + (r1 * FS + xb) * FCH + c8 * 8);
                const __half2 w00 = __float2half2_rn(ly0 * lx0), w01 = __float2half2_rn(ly0 * lx1);
                const __half2 w10 = __float2half2_rn(ly1 * lx0), w11 = __float2half2_rn(ly1 * lx1);
                const __half2 zero = __float2half2_rn(0.f);
                const __half2* ha = reinterpret_cast<const __half2*>(&ua);
                const __half2* hb = reinterpret_cast<const __half2*>(&ub);
                const __half2* hc = reinterpret_cast<const __half2*>(&uc);
                const __half2* hd = reinterpret_cast<const __half2*>(&ud);
                uint4 uo;
                __half2* ho = reinterpret_cast<__half2*>(&uo);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    ho[i] = __hmax2(__hfma2(w11, hd[i], __hfma2(w10, hc[i], __hfma2(w01, hb[i], __hmul2(w00, ha[i])))), zero);
                if (cl0 && cl0_dt == 2) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(cl0) + off) = uo;
                if (cl1 && cl1_dt == 2) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(cl1) + off) = uo;
                if ((cl0 && cl0_dt != 2) || (cl1 && cl1_dt != 2)) {
                    float r[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(ho[i]); r[2 * i] = f.x; r[2 * i + 1] = f.y; }
                    if (cl0 && cl0_dt != 2) store8_dt(cl0, off, cl0_dt, r);
                    if (cl1 && cl1_dt != 2) store8_dt(cl1, off, cl1_dt, r);
                }
            } else {
                float a[8], bb[8], c[8], d[8], r[8];
                unpack8<TZ>(s1 + (r0 * FS + xa) * FCH + c8 * 8, a);
                unpack8<TZ>(s1 + (r0 * FS + xb) * FCH + c8 * 8, bb);
                unpack8<TZ>(s1 + (r1 * FS + xa) * FCH + c8 * 8, c);
                unpack8<TZ>(s1 + (r1 * FS + xb) * FCH + c8 * 8, d);
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = fmaxf(ly0 * (lx0 * a[i] + lx1 * bb[i]) + ly1 * (lx0 * c[i] + lx1 * d[i]), 0.f);
                if (cl0) store8_dt(cl0, off, cl0_dt, r);
                if (cl1) {
                    // the TF32 store rounds r in place: keep it second when both are present (cl0 is never fp32 then)
                    store8_dt(cl1, off, cl1_dt, r);
                }
            }
        }
    }
    __syncthreads();
    // ---- NCHW: item = (channel, row, 4 consecutive x) ----
    if (out_nchw) {
        float* o = out_nchw + (int64_t)b * o_bs + (int64_t)g * o_gs + (int64_t)y0 * FO;
#pragma unroll 2
        for (int it = threadIdx.x; it < FCH * FSTRIP * (FO / 4); it += 256) {
            const int x4 = it & 15, yy = (it >> 4) & 3, c = it >> 6;
            const int r0 = tab.i0[y0 + yy] - sr0, r1 = tab.i1[y0 + yy] - sr0;
            const float ly0 = tab.l0[y0 + yy], ly1 = tab.l1[y0 + yy];
            const TZ* t0 = s2 + c * SLD + r0 * FS;
            const TZ* t1 = s2 + c * SLD + r1 * FS;
            float r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = x4 * 4 + i;
                const int xa = tab.i0[x], xb = tab.i1[x];
                const float lx0 = tab.l0[x], lx1 = tab.l1[x];
                r[i] = fmaxf(ly0 * (lx0 * ActT<TZ>::ld(t0 + xa) + lx1 * ActT<TZ>::ld(t0 + xb)) +
                             ly1 * (lx0 * ActT<TZ>::ld(t1 + xa) + lx1 * ActT<TZ>::ld(t1 + xb)), 0.f);
            }
            __stcs(reinterpret_cast<float4*>(o + ((int64_t)c * FO + yy) * FO + x4 * 4), make_float4(r[0], r[1], r[2], r[3]));
        }
    }
}

// head tail, fast: up2 + ReLU + 1x1 C->J.  One thread = one output pixel, all J outputs; the source tile is stored
// pixel-major with a 130-element pixel stride so that the per-pixel gathers of a warp never share a bank.
template <typename TZ, typename TH>
__global__ void __launch_bounds__(256)
head_up_conv_fast_kernel(const TZ* __restrict__ z, const float* __restrict__ w, const float* __restrict__ bias, int4 wsel,
                         int B, int J, float* __restrict__ hm, int64_t hm_bs, int64_t hm_gs, TH* __restrict__ hm_t) {
    constexpr int PST = 130;
    extern __shared__ __align__(16) uint8_t fsm[];
    float* sw = reinterpret_cast<float*>(fsm);             // [FCH][HJ]
    TZ* s = reinterpret_cast<TZ*>(sw + FCH * HJ);          // [FROWS*FS px][PST]
    __shared__ float sb[HJ];
    const int y0 = blockIdx.x * FSTRIP;
    const int g = blockIdx.y / B, b = blockIdx.y - g * B;
    const int img = g * B + b;
    const int sel = (g == 0) ? wsel.x : (g == 1) ? wsel.y : (g == 2) ? wsel.z : wsel.w;
    const float* wv = w + (int64_t)sel * J * FCH;
    for (int i = threadIdx.x; i < FCH * HJ; i += 256) {
        const int c = i / HJ, j = i - c * HJ;
        sw[i] = (j < J) ? wv[j * FCH + c] : 0.f;
    }
    if (threadIdx.x < HJ) sb[threadIdx.x] = (threadIdx.x < J) ? bias[(int64_t)sel * J + threadIdx.x] : 0.f;
    const int sr0 = upc(y0, FS).i0;
    const int nsr = upc(y0 + FSTRIP - 1, FS).i1 - sr0 + 1;
    {
        // 8-byte pieces: 4 bf16 or 2 floats; the padded pixel stride keeps them 4-byte aligned only -> 32-bit stores
        const uint2* src = reinterpret_cast<const uint2*>(z + ((int64_t)img * FS * FS + (int64_t)sr0 * FS) * FCH);
        constexpr int PER8 = 8 / (int)sizeof(TZ);          // elements per piece
        constexpr int PPP = FCH / PER8;                    // pieces per pixel
        uint32_t* dst = reinterpret_cast<uint32_t*>(s);
        for (int i = threadIdx.x; i < nsr * FS * PPP; i += 256) {
            const int p = i / PPP, q = i - p * PPP;
            const uint2 u = __ldg(src + i);
            const int word = (p * PST + q * PER8) * (int)sizeof(TZ) / 4;
            dst[word] = u.x; dst[word + 1] = u.y;
        }
    }
    __syncthreads();
    const int yy = threadIdx.x >> 6, x = threadIdx.x & 63;
    const int y = y0 + yy;
    const UpC cy = upc(y, FS), cx = upc(x, FS);
    const TZ* p00 = s + ((cy.i0 - sr0) * FS + cx.i0) * PST;
    const TZ* p01 = s + ((cy.i0 - sr0) * FS + cx.i1) * PST;
    const TZ* p10 = s + ((cy.i1 - sr0) * FS + cx.i0) * PST;
    const TZ* p11 = s + ((cy.i1 - sr0) * FS + cx.i1) * PST;
    float acc[HJ];
#pragma unroll
    for (int j = 0; j < HJ; ++j) acc[j] = sb[j];
#pragma unroll 4
    for (int c = 0; c < FCH; c += 2) {
        float a0, a1, b0, b1, c0, c1, d0, d1;
        if (sizeof(TZ) == 2) {
            const uint32_t ua = *reinterpret_cast<const uint32_t*>(p00 + c), ub = *reinterpret_cast<const uint32_t*>(p01 + c);
            const uint32_t uc = *reinterpret_cast<const uint32_t*>(p10 + c), ud = *reinterpret_cast<const uint32_t*>(p11 + c);
            a0 = __uint_as_float(ua << 16); a1 = __uint_as_float(ua & 0xffff0000u);
            b0 = __uint_as_float(ub << 16); b1 = __uint_as_float(ub & 0xffff0000u);
            c0 = __uint_as_float(uc << 16); c1 = __uint_as_float(uc & 0xffff0000u);
            d0 = __uint_as_float(ud << 16); d1 = __uint_as_float(ud & 0xffff0000u);
        } else {
            const float2 fa = *reinterpret_cast<const float2*>(p00 + c), fb = *reinterpret_cast<const float2*>(p01 + c);
            const float2 fc = *reinterpret_cast<const float2*>(p10 + c), fd = *reinterpret_cast<const float2*>(p11 + c);
            a0 = fa.x; a1 = fa.y; b0 = fb.x; b1 = fb.y; c0 = fc.x; c1 = fc.y; d0 = fd.x; d1 = fd.y;
        }
        const float u0 = fmaxf(cy.l0 * (cx.l0 * a0 + cx.l1 * b0) + cy.l1 * (cx.l0 * c0 + cx.l1 * d0), 0.f);
        const float u1 = fmaxf(cy.l0 * (cx.l0 * a1 + cx.l1 * b1) + cy.l1 * (cx.l0 * c1 + cx.l1 * d1), 0.f);
        const float4* w0 = reinterpret_cast<const float4*>(sw + c * HJ);
#pragma unroll
        for (int q = 0; q < HJ / 4; ++q) {
            const float4 wa = w0[q], wb = w0[q + HJ / 4];
            acc[q * 4 + 0] = fmaf(wa.x, u0, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(wa.y, u0, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(wa.z, u0, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(wa.w, u0, acc[q * 4 + 3]);
            acc[q * 4 + 0] = fmaf(wb.x, u1, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(wb.y, u1, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(wb.z, u1, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(wb.w, u1, acc[q * 4 + 3]);
        }
    }
    float* o = hm + (int64_t)b * hm_bs + (int64_t)g * hm_gs + (int64_t)y * FO + x;
#pragma unroll
    for (int j = 0; j < HJ; ++j)
        if (j < J) o[(int64_t)j * FO * FO] = acc[j];
    if (hm_t) {
        TH* ot = hm_t + ((int64_t)img * J) * FO * FO + (int64_t)y * FO + x;
#pragma unroll
        for (int j = 0; j < HJ; ++j)
            if (j < J) ActT<TH>::st(ot + (int64_t)j * FO * FO, acc[j]);
    }
}

int head_up_conv(const void* z, int z_bf16, const float* w, const float* bias, const int* wsel_host, int B, int G,
                 int Hs, int Ws, int C, int J, float* hm, int64_t hm_bs, int64_t hm_gs, void* hm_t,
                 cudaStream_t st) {
    EGR_CHECK(J <= HJ && (2 * Hs) % STRIP == 0 && G <= 4, EGR_ERR_UNSUPPORTED, "head_up_conv: J=%d Hs=%d G=%d", J, Hs, G);
    if (Hs == FS && Ws == FS && C == FCH && z_bf16 >= 2) {    // tensor-core tails: fp16 z, bf16 (2) / fp16 (3) hm_t
        if (g_opt_tail_mma) return head_tail_mma(z, w, bias, wsel_host, B, G, J, hm, hm_bs, hm_gs, hm_t, z_bf16 == 3, st);
        return head_tail_tc(z, w, bias, wsel_host, B, G, J, hm, hm_bs, hm_gs, hm_t, z_bf16 == 3, st);
    }
    EGR_CHECK(z_bf16 < 2, EGR_ERR_UNSUPPORTED, "head_up_conv: fp16 z needs the 32x32x128 geometry");
    if (Hs == FS && Ws == FS && C == FCH) {
        const size_t es = z_bf16 ? 2 : 4;
        const size_t fsmem = sizeof(float) * FCH * HJ + es * FROWS * FS * 130;
        dim3 fgrid(FO / FSTRIP, G * B);
        int4 fsel = make_int4(wsel_host[0], wsel_host[1], wsel_host[2], wsel_host[3]);
        if (z_bf16) {
            auto k = head_up_conv_fast_kernel<__nv_bfloat16, __nv_bfloat16>;
            EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            k<<<fgrid, 256, fsmem, st>>>((const __nv_bfloat16*)z, w, bias, fsel, B, J, hm, hm_bs, hm_gs, (__nv_bfloat16*)hm_t);
        } else {
            auto k = head_up_conv_fast_kernel<float, float>;
            EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            k<<<fgrid, 256, fsmem, st>>>((const float*)z, w, bias, fsel, B, J, hm, hm_bs, hm_gs, (float*)hm_t);
        }
        EGR_LAUNCHED();
        return EGR_OK;
    }
    const size_t smem = sizeof(float) * ((size_t)C * (STRIP * Ws + 1) + (size_t)C * HJ);
    dim3 grid(2 * Hs / STRIP, G * B);
    int4 wsel = make_int4(wsel_host[0], wsel_host[1], wsel_host[2], wsel_host[3]);
    if (z_bf16) {
        auto k = head_up_conv_kernel<__nv_bfloat16, __nv_bfloat16>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, st>>>((const __nv_bfloat16*)z, w, bias, wsel, B, Hs, Ws, C, J, hm, hm_bs, hm_gs,
                                   (__nv_bfloat16*)hm_t);
    } else {
        auto k = head_up_conv_kernel<float, float>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, st>>>((const float*)z, w, bias, wsel, B, Hs, Ws, C, J, hm, hm_bs, hm_gs, (float*)hm_t);
    }
    EGR_LAUNCHED();
    return EGR_OK;
}

// ---------------------------------------------------------------------------------------------
// R1 tail: relu(up2(z)) -> NCHW fp32 output + NHWC T copy
// ---------------------------------------------------------------------------------------------
template <typename TZ>
__global__ void __launch_bounds__(256)
up2_relu_dual_kernel(const TZ* __restrict__ z, int B, int Hs, int Ws, int C, float* __restrict__ out_nchw,
                     int64_t o_bs, int64_t o_gs, TZ* __restrict__ out_nhwc) {
    extern __shared__ float smem[];
    const int ld = STRIP * Ws + 1;
    float* s = smem;
    const int y0 = blockIdx.x * STRIP;
    const int g = blockIdx.y / B, b = blockIdx.y - g * B;
    const int img = g * B + b;
    const int H = 2 * Hs, W = 2 * Ws;
    const int sr0 = upc(y0, Hs).i0;
    const int sr1 = upc(min(y0 + STRIP - 1, H - 1), Hs).i1;
    load_src_rows<TZ>(z + (int64_t)img * Hs * Ws * C, s, ld, sr0, sr1 - sr0 + 1, Ws, C);
    __syncthreads();
    // pass 1: NCHW, x fastest
    if (out_nchw) {
        float* o = out_nchw + (int64_t)b * o_bs + (int64_t)g * o_gs;
        for (int i = threadIdx.x; i < C * STRIP * W; i += blockDim.x) {
            const int c = i / (STRIP * W), r = i - c * (STRIP * W);
            const int yy = r / W, x = r - yy * W;
            const int y = y0 + yy;
            const float u = fmaxf(interp(s + c * ld, Ws, sr0, upc(y, Hs), upc(x, Ws)), 0.f);
            __stcs(o + ((int64_t)c * H + y) * W + x, u);
        }
    }
    // pass 2: NHWC, c fastest
    if (out_nhwc) {
        TZ* o = out_nhwc + (int64_t)img * H * W * C;
        for (int i = threadIdx.x; i < C * STRIP * W; i += blockDim.x) {
            const int px = i / C, c = i - px * C;
            const int yy = px / W, x = px - yy * W;
            const int y = y0 + yy;
            const float u = fmaxf(interp(s + c * ld, Ws, sr0, upc(y, Hs), upc(x, Ws)), 0.f);
            ActT<TZ>::st(o + ((int64_t)y * W + x) * C + c, u);
        }
    }
}

int up2_relu_dual(const void* z, int z_dt, int B, int G, int Hs, int Ws, int C, float* out_nchw, int64_t o_bs,
                  int64_t o_gs, void* cl0, int cl0_dt, void* cl1, int cl1_dt, cudaStream_t st, int fp32_interp, bool allow_mma) {
    EGR_CHECK((2 * Hs) % STRIP == 0, EGR_ERR_UNSUPPORTED, "up2_relu_dual: geometry");
    if (Hs == FS && Ws == FS && C == FCH) {
        const size_t es = z_dt ? 2 : 4;
        // the transposed tile only serves the NCHW pass: without it the block needs half the shared memory
        const size_t fsmem = es * (FROWS * FS * FCH + (out_nchw ? FCH * 130 : 0));
        // only the 16-bit channels-last map is wanted (chained forward, refined features not materialised): the bilinear
        // upsample as a tcgen05.mma (head_tail_mma.cu), the blended tile leaves as bulk tensor stores
        if (allow_mma && g_opt_tail_mma && z_dt == 2 && !out_nchw && !cl1 && cl0 && (cl0_dt == 1 || cl0_dt == 2) &&
            ((uintptr_t)z % 16) == 0 && ((uintptr_t)cl0 % 16) == 0)
            return up2_relu_mma(z, G * B, cl0, cl0_dt == 2, st);
        dim3 fgrid(FO / FSTRIP, G * B);
        if (z_dt == 2) {
            auto k = up2_relu_dual_fast_kernel<__half>;
            EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            EGR_LAUNCH(k, fgrid, 256, fsmem, st, (const __half*)z, B, out_nchw, o_bs, o_gs, cl0, cl0_dt, cl1, cl1_dt, fp32_interp);
        } else if (z_dt == 1) {
            auto k = up2_relu_dual_fast_kernel<__nv_bfloat16>;
            EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            EGR_LAUNCH(k, fgrid, 256, fsmem, st, (const __nv_bfloat16*)z, B, out_nchw, o_bs, o_gs, cl0, cl0_dt, cl1, cl1_dt, fp32_interp);
        } else {
            auto k = up2_relu_dual_fast_kernel<float>;
            EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            EGR_LAUNCH(k, fgrid, 256, fsmem, st, (const float*)z, B, out_nchw, o_bs, o_gs, cl0, cl0_dt, cl1, cl1_dt, fp32_interp);
        }
        return EGR_OK;
    }
    // general geometry (not on the shipped hot path): fp32 / bf16 z, one channels-last copy in z's type
    void* out_nhwc = cl0;
    void* out_nhwc_hp = cl1;
    const int z_bf16 = z_dt;
    EGR_CHECK(z_dt != 2 && (!cl0 || cl0_dt == (z_dt ? 1 : 3)), EGR_ERR_UNSUPPORTED, "up2_relu_dual: general geometry takes fp32 / bf16 z with a same-typed copy");
    EGR_CHECK(!out_nhwc_hp, EGR_ERR_UNSUPPORTED, "up2_relu_dual: the high-precision channels-last export needs the 32x32x128 geometry");
    const size_t smem = sizeof(float) * (size_t)C * (STRIP * Ws + 1);
    dim3 grid(2 * Hs / STRIP, G * B);
    if (z_bf16) {
        auto k = up2_relu_dual_kernel<__nv_bfloat16>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, st>>>((const __nv_bfloat16*)z, B, Hs, Ws, C, out_nchw, o_bs, o_gs, (__nv_bfloat16*)out_nhwc);
    } else {
        auto k = up2_relu_dual_kernel<float>;
        EGR_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, st>>>((const float*)z, B, Hs, Ws, C, out_nchw, o_bs, o_gs, (float*)out_nhwc);
    }
    EGR_LAUNCHED();
    return EGR_OK;
}

// ---------------------------------------------------------------------------------------------
// 16 bytes (4 floats / 8 bf16) of channels per thread
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) {
        uint4 r;
        r.x = __float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x)));
        r.y = __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y)));
        r.z = __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z)));
        r.w = __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w)));
        return r;
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ uint32_t m2(uint32_t a, uint32_t b) {
        const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
        return *reinterpret_cast<const uint32_t*>(&r);
    }
    static __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) { return make_uint4(m2(a.x, b.x), m2(a.y, b.y), m2(a.z, b.z), m2(a.w, b.w)); }
};

template <> struct Vec16<__half> {
    static constexpr int N = 8;
    static __device__ __forceinline__ uint32_t m2(uint32_t a, uint32_t b) {
        const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
        return *reinterpret_cast<const uint32_t*>(&r);
    }
    static __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) { return make_uint4(m2(a.x, b.x), m2(a.y, b.y), m2(a.z, b.z), m2(a.w, b.w)); }
};

template <typename T>
__global__ void maxpool2_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t total_vec, int H, int W, int C) {
    constexpr int VN = Vec16<T>::N;
    pdl_trigger();
    pdl_wait();
    const int Ho = H >> 1, Wo = W >> 1, CV = C / VN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        int64_t r = i / CV;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int64_t img = r / Ho;
        const uint4* p = reinterpret_cast<const uint4*>(in + ((img * H + 2 * oy) * W + 2 * ox) * C) + cv;
        const uint4 a = __ldg(p), b = __ldg(p + CV), c = __ldg(p + (int64_t)W * CV), d = __ldg(p + (int64_t)W * CV + CV);
        reinterpret_cast<uint4*>(out)[i] = Vec16<T>::vmax(Vec16<T>::vmax(a, b), Vec16<T>::vmax(c, d));
    }
}

int maxpool2_nhwc(const void* in, void* out, int dt, int64_t n_img, int H, int W, int C, cudaStream_t st) {
    const int is_bf16 = (dt == 1);
    const int vn = dt ? 8 : 4;
    EGR_CHECK(C % vn == 0, EGR_ERR_UNSUPPORTED, "maxpool2: C=%d", C);
    const int64_t total = n_img * (H / 2) * (W / 2) * (C / vn);
    if (total == 0) return EGR_OK;
    const int grid = (int)(ceil_div64(total, 256) < 148 * 32 ? ceil_div64(total, 256) : 148 * 32);
    if (dt == 2) EGR_LAUNCH(maxpool2_kernel<__half>, grid, 256, 0, st, (const __half*)in, (__half*)out, total, H, W, C);
    else if (is_bf16) EGR_LAUNCH(maxpool2_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, total, H, W, C);
    else EGR_LAUNCH(maxpool2_kernel<float>, grid, 256, 0, st, (const float*)in, (float*)out, total, H, W, C);
    return EGR_OK;
}

// ---------------------------------------------------------------------------------------------
// prepack helpers (run once per weight load; simple grid-stride kernels)
// ---------------------------------------------------------------------------------------------
__global__ void repack_conv3_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin) {
    const int64_t total = (int64_t)Cout * Cin * 9;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // out index: ((co*3 + ky)*3 + kx)*Cin + ci
        const int ci = (int)(i % Cin);
        int64_t r = i / Cin;
        const int kx = (int)(r % 3); r /= 3;
        const int ky = (int)(r % 3);
        const int64_t co = r / 3;
        out[i] = w[((co * Cin + ci) * 3 + ky) * 3 + kx];
    }
}
int repack_conv3(const float* w, float* out, int Cout, int Cin, cudaStream_t st) {
    repack_conv3_kernel<<<(int)ceil_div64((int64_t)Cout * Cin * 9, 256), 256, 0, st>>>(w, out, Cout, Cin);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void round_tf32_kernel(float* __restrict__ p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = round_tf32(p[i]);
}
int round_tf32_inplace(float* p, int64_t n, cudaStream_t st) {
    if (n == 0) return EGR_OK;
    const int64_t g = ceil_div64(n, 256);
    round_tf32_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(p, n);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        ActT<__half>::st(out + i, in[i]);
}
int cast_f16(const float* in, __half* out, int64_t n, cudaStream_t st) {
    if (n == 0) return EGR_OK;
    const int64_t g = ceil_div64(n, 256);
    cast_f16_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(in, out, n);
    EGR_LAUNCHED();
    return EGR_OK;
}

int cast_bf16(const float* in, __nv_bfloat16* out, int64_t n, cudaStream_t st) {
    if (n == 0) return EGR_OK;
    const int64_t g = ceil_div64(n, 256);
    cast_bf16_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(in, out, n);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void split_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n, int K) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K;
        const int k = (int)(i - r * K);
        const float w = in[i];
        __half hi, lo;
        ActT<__half>::st(&hi, w);
        ActT<__half>::st(&lo, w - __half2float(hi));
        out[r * 2 * K + k] = hi;
        out[r * 2 * K + K + k] = lo;
    }
}
int split_f16(const float* in, __half* out, int64_t R, int K, cudaStream_t st) {
    const int64_t n = R * K;
    if (n == 0) return EGR_OK;
    const int64_t g = ceil_div64(n, 256);
    split_f16_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(in, out, n, K);
    EGR_LAUNCHED();
    return EGR_OK;
}
__global__ void split3_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, int K) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K;
        const int k = (int)(i - r * K);
        const float w = in[i];
        const float hi = round_tf32(w);
        out[r * 3 * K + k] = hi;
        out[r * 3 * K + K + k] = hi;
        out[r * 3 * K + 2 * K + k] = round_tf32(w - hi);
    }
}
int split3_tf32(const float* in, float* out, int64_t R, int K, cudaStream_t st) {
    const int64_t n = R * K;
    if (n == 0) return EGR_OK;
    const int64_t g = ceil_div64(n, 256);
    split3_tf32_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(in, out, n, K);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void split3_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n, int K) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K;
        const int k = (int)(i - r * K);
        const float w = in[i];
        const __half hi = __float2half_rn(w);
        out[r * 3 * K + k] = hi;
        out[r * 3 * K + K + k] = hi;
        out[r * 3 * K + 2 * K + k] = __float2half_rn(w - __half2float(hi));
    }
}
int split3_f16(const float* in, __half* out, int64_t R, int K, cudaStream_t st) {
    const int64_t n = R * K;
    if (n == 0) return EGR_OK;
    const int64_t g = ceil_div64(n, 256);
    split3_f16_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(in, out, n, K);
    EGR_LAUNCHED();
    return EGR_OK;
}

int cast_act(const float* in, void* out, int out_bf16, int64_t n, cudaStream_t st) {
    if (out_bf16 == 2) return cast_f16(in, (__half*)out, n, st);
    if (out_bf16) return cast_bf16(in, (__nv_bfloat16*)out, n, st);
    if (n == 0) return EGR_OK;
    const int64_t g = ceil_div64(n, 256);
    copy_f32_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(in, (float*)out, n);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void transpose2d_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        t[i][threadIdx.x] = (r < R && c < C) ? in[(int64_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) out[(int64_t)c * R + r] = t[threadIdx.x][i];
    }
}
int transpose2d(const float* in, float* out, int R, int C, cudaStream_t st) {
    dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
    transpose2d_kernel<<<grid, block, 0, st>>>(in, out, R, C);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void small_matmul_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                    const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N;
        const int n = (int)(i - m * N);
        float acc = bias ? bias[n] : 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(A[m * K + k], B[(int64_t)k * N + n], acc);
        C[i] = acc;
    }
}
int small_matmul(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, cudaStream_t st) {
    const int64_t g = ceil_div64((int64_t)M * N, 256);
    small_matmul_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(A, B, bias, C, M, N, K);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void add_rowvec_kernel(const float* __restrict__ in, const float* __restrict__ v, float* __restrict__ out,
                                  int64_t total, int C) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i] + v[i % C];
}
int add_rowvec(const float* in, const float* v, float* out, int64_t R, int C, cudaStream_t st) {
    const int64_t total = R * C, g = ceil_div64(total, 256);
    add_rowvec_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(in, v, out, total, C);
    EGR_LAUNCHED();
    return EGR_OK;
}

__global__ void permute_mlp_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int64_t total, int V,
                                          int C, int HWp) {
    const int K = V * C * HWp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // out k' = (v*HWp + p)*C + c   <-   in k = (v*C + c)*HWp + p
        const int kp = (int)(i % K);
        const int64_t n = i / K;
        const int c = kp % C;
        const int r = kp / C;
        const int p = r % HWp, v = r / HWp;
        out[i] = w[n * K + ((int64_t)(v * C + c)) * HWp + p];
    }
}
int permute_mlp_weight(const float* w, float* out, int N, int V, int C, int HWp, cudaStream_t st) {
    const int64_t total = (int64_t)N * V * C * HWp, g = ceil_div64(total, 256);
    permute_mlp_weight_kernel<<<(int)(g < 148 * 64 ? g : 148 * 64), 256, 0, st>>>(w, out, total, V, C, HWp);
    EGR_LAUNCHED();
    return EGR_OK;
}

}  // namespace egr
