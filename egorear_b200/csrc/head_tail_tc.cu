// Heatmap-head tail on the tensor cores (EGR_PREC_BF16, shipped geometry 32x32 -> 64x64, C = 128; z in fp16):
//
//   hm[j][y][x] = sum_c W[j][c] * relu(up2_bilinear_align_corners(z)[y][x][c]) + b[j]        j < 15
//
// (the `nn.Upsample, ReLU, Conv2d(128, 15, 1)` end of conv_heatmap_layers*, egoposeformer_heatmap_mvf_ex.py:108-110,
// :579-583, with the preceding 1x1 conv commuted in front of the upsample).  The SIMT version spends 16 FMAs per
// interpolated value on the 1x1 conv; here one CTA = 4 output rows = 256 pixels = two M=128 tcgen05 tiles:
//   1. the (at most) 4 source rows of z are staged in shared memory (padded pixel stride, conflict-free LDS.128)
//   2. thread t interpolates pixel t (64 channels = one k-block at a time) and applies ReLU in half2 arithmetic (z is
//      fp16: two channels per HFMA2, no unpacking) and writes the fp16 row straight into the K-major SWIZZLE_128B A tile
//   3. one thread issues 2 x 4 tcgen05.mma (M=128, N=16, K=16) per k-block against the bf16 weight tile, accumulators
//      in TMEM; 71 KB of shared memory -> three CTAs per SM overlap each other's load / MMA / store phases
//   4. tcgen05.ld gives thread t the 16 joint values of pixel t; 32 lanes = 32 consecutive x -> 128 B stores per joint
#include "layout_ops.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <type_traits>

namespace egr {
using namespace tcx;
namespace {

constexpr int HT_FS = 32, HT_FO = 64, HT_C = 128, HT_STRIP = 4, HT_ROWS = 4, HT_NJ = 16;
constexpr int HT_PST = 272;                                  // bytes per staged source pixel: 256 + 16
constexpr int HT_OFF_A = 0;                                  // [2 M-tiles][128 rows][128 B] = 32 KB (one 64-channel k-block at a time)
constexpr int HT_OFF_W = 2 * 128 * 128;                      // [2 k-blocks][16 rows][128 B] =  4 KB (+ 4 KB W_lo when PRECISE)
// PRECISE (EGR_PREC_FP16): the 1x1 weights as an fp16 pair W_hi + W_lo (W_lo = fp16 of the rounding residual), twice the
// MMAs against the same A tile: the weight rounding error of this last layer disappears; hm_t is written in fp16
template <bool PRECISE> struct HtCfg {
    static constexpr int W_BYTES = (PRECISE ? 2 : 1) * 2 * 16 * 128;
    static constexpr int OFF_SRC = HT_OFF_W + W_BYTES;       // [4 rows x 32 px][272 B]
    static constexpr int OFF_BAR = OFF_SRC + HT_ROWS * HT_FS * HT_PST;
    static constexpr int SMEM = OFF_BAR + 64 + 1024 /*align*/;
};
constexpr uint32_t HT_TMEM_COLS = 32;

struct HtUp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ HtUp ht_up(int dst) {
    // nn.Upsample(scale_factor=2, bilinear, align_corners=True): src = dst * (in-1)/(out-1)
    const float scale = (float)(HT_FS - 1) / (float)(HT_FO - 1);
    const float s = scale * (float)dst;
    HtUp c;
    c.i0 = (int)s;
    c.i1 = c.i0 + ((c.i0 < HT_FS - 1) ? 1 : 0);
    c.l1 = s - (float)c.i0;
    c.l0 = 1.f - c.l1;
    return c;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// relu(w00*a + w01*b + w10*c + w11*d) on one fp16x2 word
__device__ __forceinline__ uint32_t interp_word(uint32_t a, uint32_t b, uint32_t c, uint32_t d, __half2 w00, __half2 w01,
                                                __half2 w10, __half2 w11) {
    const __half2 r = __hmax2(__hfma2(w11, *reinterpret_cast<const __half2*>(&d),
                              __hfma2(w10, *reinterpret_cast<const __half2*>(&c),
                              __hfma2(w01, *reinterpret_cast<const __half2*>(&b), __hmul2(w00, *reinterpret_cast<const __half2*>(&a))))),
                              __float2half2_rn(0.f));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// PRECISE: the same in fp32 arithmetic, one rounding (of the result) instead of one per HFMA2 plus the weights'
__device__ __forceinline__ uint32_t interp_word_f32(uint32_t a, uint32_t b, uint32_t c, uint32_t d, float w00, float w01,
                                                    float w10, float w11) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a)), fb = __half22float2(*reinterpret_cast<const __half2*>(&b));
    const float2 fc = __half22float2(*reinterpret_cast<const __half2*>(&c)), fd = __half22float2(*reinterpret_cast<const __half2*>(&d));
    const float lo = fmaxf(fmaf(w11, fd.x, fmaf(w10, fc.x, fmaf(w01, fb.x, w00 * fa.x))), 0.f);
    const float hi = fmaxf(fmaf(w11, fd.y, fmaf(w10, fc.y, fmaf(w01, fb.y, w00 * fa.y))), 0.f);
    return pack_f16x2(lo, hi);
}

template <bool PRECISE>
__global__ void __launch_bounds__(256, 3)
head_tail_tc_kernel(const __half* __restrict__ z, const float* __restrict__ w, const float* __restrict__ bias, int4 wsel,
                    int B, int J, float* __restrict__ hm, int64_t hm_bs, int64_t hm_gs, void* __restrict__ hm_t_) {
    using HM_T = typename std::conditional<PRECISE, __half, __nv_bfloat16>::type;
    HM_T* __restrict__ hm_t = reinterpret_cast<HM_T*>(hm_t_);
    constexpr int HT_OFF_SRC = HtCfg<PRECISE>::OFF_SRC, HT_OFF_BAR = HtCfg<PRECISE>::OFF_BAR;
    extern __shared__ __align__(1024) uint8_t ht_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ht_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem + HT_OFF_A;
    uint8_t* sW = smem + HT_OFF_W;
    uint8_t* sS = smem + HT_OFF_SRC;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + HT_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    float* sb = reinterpret_cast<float*>(bar + 2);           // [16]
    const int tid = threadIdx.x, warp = tid >> 5;
    const int y0 = blockIdx.x * HT_STRIP;
    const int g = blockIdx.y / B, b = blockIdx.y - g * B;
    const int img = g * B + b;
    const int sel = (g == 0) ? wsel.x : (g == 1) ? wsel.y : (g == 2) ? wsel.z : wsel.w;

    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), HT_TMEM_COLS);
        pdl_trigger();          // only after this CTA owns its TMEM columns (common.cuh)
    }
    pdl_wait();
    if (tid < HT_NJ) sb[tid] = (tid < J) ? __ldg(bias + (int64_t)sel * J + tid) : 0.f;
    {   // weight tile: fp32 [J][128] -> fp16 [2 k-blocks][16 rows][64], rows >= J zero
        const int n = tid >> 4, piece = tid & 15;
        uint4 u = make_uint4(0u, 0u, 0u, 0u), ul = make_uint4(0u, 0u, 0u, 0u);
        if (n < J) {
            const float4* src = reinterpret_cast<const float4*>(w + ((int64_t)sel * J + n) * HT_C + piece * 8);
            const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
            u = make_uint4(pack_f16x2(f0.x, f0.y), pack_f16x2(f0.z, f0.w), pack_f16x2(f1.x, f1.y), pack_f16x2(f1.z, f1.w));
            if (PRECISE) {      // residual of the fp16 rounding, itself in fp16
                const __half2* h = reinterpret_cast<const __half2*>(&u);
                const float2 a = __half22float2(h[0]), b2 = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
                ul = make_uint4(pack_f16x2(f0.x - a.x, f0.y - a.y), pack_f16x2(f0.z - b2.x, f0.w - b2.y),
                                pack_f16x2(f1.x - c.x, f1.y - c.y), pack_f16x2(f1.z - d.x, f1.w - d.y));
            }
        }
        const int kb = piece >> 3, pp = piece & 7;
        *reinterpret_cast<uint4*>(sW + kb * 2048 + n * 128 + ((pp ^ (n & 7)) << 4)) = u;
        if (PRECISE) *reinterpret_cast<uint4*>(sW + 4096 + kb * 2048 + n * 128 + ((pp ^ (n & 7)) << 4)) = ul;
    }
    const int sr0 = ht_up(y0).i0;
    const int nsr = ht_up(y0 + HT_STRIP - 1).i1 - sr0 + 1;
    {   // source rows [sr0, sr0 + nsr) of this image
        const uint4* src = reinterpret_cast<const uint4*>(z + ((int64_t)img * HT_FS * HT_FS + (int64_t)sr0 * HT_FS) * HT_C);
        const uint32_t sS32 = smem_u32(sS);
        for (int i = tid; i < nsr * HT_FS * 16; i += 256) {
            const int px = i >> 4, pc = i & 15;
            cp_async16(sS32 + px * HT_PST + pc * 16, src + i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- interpolate + ReLU: pixel `tid` -> row (tid & 127) of A tile (tid >> 7), one 64-channel k-block at a time ----
    // (the A tiles hold a single k-block so that three CTAs fit an SM; the second half is interpolated once the MMAs of the
    //  first have finished reading the tile - the other resident CTAs fill that wait)
    const int yy = tid >> 6, x = tid & 63;
    const int y = y0 + yy;
    {
        const HtUp cy = ht_up(y), cx = ht_up(x);
        const uint32_t sS32 = smem_u32(sS);
        const uint32_t p00 = sS32 + ((cy.i0 - sr0) * HT_FS + cx.i0) * HT_PST;
        const uint32_t p01 = sS32 + ((cy.i0 - sr0) * HT_FS + cx.i1) * HT_PST;
        const uint32_t p10 = sS32 + ((cy.i1 - sr0) * HT_FS + cx.i0) * HT_PST;
        const uint32_t p11 = sS32 + ((cy.i1 - sr0) * HT_FS + cx.i1) * HT_PST;
        const float f00 = cy.l0 * cx.l0, f01 = cy.l0 * cx.l1, f10 = cy.l1 * cx.l0, f11 = cy.l1 * cx.l1;
        const __half2 w00 = __float2half2_rn(f00), w01 = __float2half2_rn(f01);
        const __half2 w10 = __float2half2_rn(f10), w11 = __float2half2_rn(f11);
        const int r = tid & 127;
        const uint32_t arow = smem_u32(sA) + (tid >> 7) * 16384 + r * 128;
        const int sw = r & 7;
#pragma unroll 1
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const int pc = kb * 8 + pp;
                const uint4 a = lds128(p00 + pc * 16), bb = lds128(p01 + pc * 16);
                const uint4 c = lds128(p10 + pc * 16), d = lds128(p11 + pc * 16);
                uint4 o;
                if (PRECISE) {
                    o.x = interp_word_f32(a.x, bb.x, c.x, d.x, f00, f01, f10, f11);
                    o.y = interp_word_f32(a.y, bb.y, c.y, d.y, f00, f01, f10, f11);
                    o.z = interp_word_f32(a.z, bb.z, c.z, d.z, f00, f01, f10, f11);
                    o.w = interp_word_f32(a.w, bb.w, c.w, d.w, f00, f01, f10, f11);
                } else {
                    o.x = interp_word(a.x, bb.x, c.x, d.x, w00, w01, w10, w11);
                    o.y = interp_word(a.y, bb.y, c.y, d.y, w00, w01, w10, w11);
                    o.z = interp_word(a.z, bb.z, c.z, d.z, w00, w01, w10, w11);
                    o.w = interp_word(a.w, bb.w, c.w, d.w, w00, w01, w10, w11);
                }
                sts128(arow + ((pp ^ sw) << 4), o);
            }
            fence_async_smem();            // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                constexpr uint32_t idesc = make_idesc_fmt(128, HT_NJ, 0u);      // fp16 operands
                const uint64_t db = make_smem_desc(smem_u32(sW + kb * 2048));
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint64_t da = make_smem_desc(smem_u32(sA + mt * 16384));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma<false>(tmem_base + mt * HT_NJ, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
                    if (PRECISE) {
                        const uint64_t dl = make_smem_desc(smem_u32(sW + 4096 + kb * 2048));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) tc_mma<false>(tmem_base + mt * HT_NJ, da + 2 * kk, dl + 2 * kk, idesc, 1u);
                    }
                }
                tc_commit(smem_u32(bar));
            }
            __syncwarp();
            mbar_wait(smem_u32(bar), kb);      // k-block 0: the tile may be overwritten; k-block 1: accumulators complete
        }
    }
    tc_fence_after();
    uint32_t v[16];
    tc_ld16_issue(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * HT_NJ, v);
    tc_ld16_wait(v);
    // thread `tid` now holds the joints of pixel `tid`; lanes = consecutive x
    float r[HT_NJ];
#pragma unroll
    for (int j4 = 0; j4 < HT_NJ; j4 += 4) {
        const uint4 bq = lds128(smem_u32(sb + j4));
        r[j4 + 0] = __uint_as_float(v[j4 + 0]) + __uint_as_float(bq.x); r[j4 + 1] = __uint_as_float(v[j4 + 1]) + __uint_as_float(bq.y);
        r[j4 + 2] = __uint_as_float(v[j4 + 2]) + __uint_as_float(bq.z); r[j4 + 3] = __uint_as_float(v[j4 + 3]) + __uint_as_float(bq.w);
    }
    float* o = hm + (int64_t)b * hm_bs + (int64_t)g * hm_gs + (int64_t)y * HT_FO + x;
#pragma unroll
    for (int j = 0; j < HT_NJ; ++j)
        if (j < J) o[(int64_t)j * HT_FO * HT_FO] = r[j];
    if (hm_t) {
        HM_T* ot = hm_t + ((int64_t)img * J) * HT_FO * HT_FO + (int64_t)y * HT_FO + x;
#pragma unroll
        for (int j = 0; j < HT_NJ; ++j)
            if (j < J) ActT<HM_T>::st(ot + (int64_t)j * HT_FO * HT_FO, r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, HT_TMEM_COLS);
    }
}

}  // namespace

int head_tail_tc(const void* z, const float* w, const float* bias, const int* wsel_host, int B, int G, int J, float* hm,
                 int64_t hm_bs, int64_t hm_gs, void* hm_t, int precise, cudaStream_t st) {
    EGR_CHECK(J <= HT_NJ && G <= 4, EGR_ERR_UNSUPPORTED, "head_tail_tc: J=%d G=%d", J, G);
    static bool attr_set = false;
    if (!attr_set) {
        EGR_CUDA_OK(cudaFuncSetAttribute(head_tail_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HtCfg<false>::SMEM));
        EGR_CUDA_OK(cudaFuncSetAttribute(head_tail_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HtCfg<true>::SMEM));
        attr_set = true;
    }
    dim3 grid(HT_FO / HT_STRIP, G * B);
    const int4 sel = make_int4(wsel_host[0], wsel_host[1], wsel_host[2], wsel_host[3]);
    if (precise)
        EGR_LAUNCH(head_tail_tc_kernel<true>, grid, 256, HtCfg<true>::SMEM, st, (const __half*)z, w, bias, sel, B, J, hm, hm_bs, hm_gs, hm_t);
    else
        EGR_LAUNCH(head_tail_tc_kernel<false>, grid, 256, HtCfg<false>::SMEM, st, (const __half*)z, w, bias, sel, B, J, hm, hm_bs, hm_gs, hm_t);
    return EGR_OK;
}

}  // namespace egr
