// mvfex token -> feature head on the tensor cores (EGR_PREC_BF16):  post_norm LayerNorm, the [15 x 256] token matrix
// re-read as a 15-channel 16x16 image (egoposeformer_heatmap_mvf_ex.py:707-711), TransformerHeadLayer's 1x1(15->64) +
// ReLU, and its 1x1(64->128) commuted in front of the bilinear x2 (:938-970).  One CTA per (refiner, frame):
//   1. 8 warps LayerNorm the 15 token rows into shared memory (fp32)
//   2. thread p owns image position p: 64 hidden values = relu(W0 . x[:, p] + b0) in fp32 (weights are smem broadcasts),
//      packed to bf16 straight into the K-major SWIZZLE_128B A tile (K = 64 = one 128-byte row)
//   3. one thread issues 2 x 4 tcgen05.mma (M=128, N=128, K=16) against the bf16 W3 tile; accumulators in TMEM
//   4. tcgen05.ld returns position p's 128 outputs to thread p: + bias, bf16, 256 contiguous bytes per thread
// Replaces tok_ln_image_kernel + two SIMT GEMM launches (6.6 + 11 + 38 us under ncu).
#include "token_batched.cuh"
#include "tc_ptx.cuh"

namespace egr {
using namespace tcx;
namespace {

constexpr int TH_E = 256, TH_J = 16, TH_HID = 64, TH_OUT = 128;
constexpr int TH_OFF_A = 0;                                  // [2 M-tiles][128 rows][128 B]  = 32 KB
constexpr int TH_OFF_B = 2 * 128 * 128;                      // [128 rows][128 B]             = 16 KB
constexpr int TH_OFF_YS = TH_OFF_B + 128 * 128;              // [16][256] fp32                = 16 KB
constexpr int TH_OFF_W0 = TH_OFF_YS + TH_J * TH_E * 4;       // [64][16] fp32                 =  4 KB
constexpr int TH_OFF_B0 = TH_OFF_W0 + TH_HID * TH_J * 4;     // [64] fp32
constexpr int TH_OFF_B3 = TH_OFF_B0 + TH_HID * 4;            // [128] fp32
constexpr int TH_OFF_BAR = TH_OFF_B3 + TH_OUT * 4;
constexpr int TH_SMEM = TH_OFF_BAR + 64 + 1024 /*align*/;
// fp16 variant (EGR_PREC_FP16): a second B tile holds W3_lo, the fp16 rounding residual of W3 (twice the MMAs, no weight error)
constexpr int TH_OFF_BLO = (TH_OFF_BAR + 64 + 1023) / 1024 * 1024;
constexpr int TH_SMEM_F16 = TH_OFF_BLO + 128 * 128 + 1024 /*align*/;
constexpr uint32_t TH_TMEM_COLS = 256;

template <bool F16> __device__ __forceinline__ uint32_t th_pack(float lo, float hi) {
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool F16> __device__ __forceinline__ uint32_t th_pack_relu(float lo, float hi) {
    uint32_t r;
    if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// fp16 pair -> the fp32 values it holds
__device__ __forceinline__ float2 th_unpack_f16(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

template <bool F16>
__global__ void __launch_bounds__(256, 2)
tok_head_tc_kernel(const float* __restrict__ x, int ldx, int B, int J, const float* const* __restrict__ gamma,
                   const float* const* __restrict__ beta, const float* __restrict__ w0, const float* __restrict__ b0,
                   const float* __restrict__ w3, const float* __restrict__ b3, int r0, uint16_t* __restrict__ t1) {
    extern __shared__ __align__(1024) uint8_t th_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(th_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem + TH_OFF_A;
    uint8_t* sB = smem + TH_OFF_B;
    float* ys = reinterpret_cast<float*>(smem + TH_OFF_YS);
    float* sW0 = reinterpret_cast<float*>(smem + TH_OFF_W0);
    float* sb0 = reinterpret_cast<float*>(smem + TH_OFF_B0);
    float* sb3 = reinterpret_cast<float*>(smem + TH_OFF_B3);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TH_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x / B;
    const int set = r0 + g;

    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), TH_TMEM_COLS);
        pdl_trigger();          // only after this CTA owns its TMEM columns (common.cuh)
    }
    pdl_wait();
    // ---- weights -> smem ----
    for (int i = tid; i < TH_HID * TH_J; i += 256) sW0[i] = __ldg(w0 + (int64_t)set * TH_HID * TH_J + i);
    if (tid < TH_HID) sb0[tid] = __ldg(b0 + set * TH_HID + tid);
    if (tid < TH_OUT) sb3[tid] = __ldg(b3 + set * TH_OUT + tid);
    {   // W3 fp32 [128][64] -> bf16 K-major SWIZZLE_128B tile: row n = 128 B, 8 pieces of 8 channels
        const int n = tid >> 1, h4 = (tid & 1) * 4;
        const float4* src = reinterpret_cast<const float4*>(w3 + ((int64_t)set * TH_OUT + n) * TH_HID) + h4 * 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 f0 = __ldg(src + 2 * q), f1 = __ldg(src + 2 * q + 1);
            const uint4 u = make_uint4(th_pack<F16>(f0.x, f0.y), th_pack<F16>(f0.z, f0.w), th_pack<F16>(f1.x, f1.y), th_pack<F16>(f1.z, f1.w));
            *reinterpret_cast<uint4*>(sB + n * 128 + (((h4 + q) ^ (n & 7)) << 4)) = u;
            if (F16) {
                const float2 a = th_unpack_f16(u.x), b2 = th_unpack_f16(u.y), c = th_unpack_f16(u.z), d = th_unpack_f16(u.w);
                const uint4 ul = make_uint4(th_pack<true>(f0.x - a.x, f0.y - a.y), th_pack<true>(f0.z - b2.x, f0.w - b2.y),
                                            th_pack<true>(f1.x - c.x, f1.y - c.y), th_pack<true>(f1.z - d.x, f1.w - d.y));
                *reinterpret_cast<uint4*>(smem + TH_OFF_BLO + n * 128 + (((h4 + q) ^ (n & 7)) << 4)) = ul;
            }
        }
    }
    // ---- post_norm LayerNorm of the J token rows (eps 1e-5, biased variance), row J..15 = 0 ----
    {
        const float* xin = x + (int64_t)blockIdx.x * J * ldx;
        const float* ga = gamma[g];
        const float* be = beta[g];
        for (int j = warp; j < TH_J; j += 8) {
            if (j < J) {
                float v[TH_E / 32];
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < TH_E / 32; ++i) { v[i] = xin[(int64_t)j * ldx + lane + 32 * i]; s += v[i]; }
                const float mean = warp_sum(s) * (1.f / TH_E);
                float q = 0.f;
#pragma unroll
                for (int i = 0; i < TH_E / 32; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
                const float rstd = rsqrtf(warp_sum(q) * (1.f / TH_E) + 1e-5f);
#pragma unroll
                for (int i = 0; i < TH_E / 32; ++i) {
                    const int c = lane + 32 * i;
                    ys[j * TH_E + c] = (v[i] - mean) * rstd * __ldg(ga + c) + __ldg(be + c);
                }
            } else {
#pragma unroll
                for (int i = 0; i < TH_E / 32; ++i) ys[j * TH_E + lane + 32 * i] = 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- thread = image position: hidden = relu(W0 . x + b0), bf16, into the A tile ----
    {
        float xv[TH_J];
#pragma unroll
        for (int j = 0; j < TH_J; ++j) xv[j] = ys[j * TH_E + tid];
        const int r = tid & 127;
        uint8_t* arow = sA + (tid >> 7) * 16384 + r * 128;
#pragma unroll 2
        for (int n8 = 0; n8 < TH_HID; n8 += 8) {
            float hacc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4* wr = reinterpret_cast<const float4*>(sW0 + (n8 + i) * TH_J);
                float a = sb0[n8 + i];
#pragma unroll
                for (int q = 0; q < TH_J / 4; ++q) {
                    const float4 ww = wr[q];
                    a = fmaf(ww.x, xv[4 * q + 0], a); a = fmaf(ww.y, xv[4 * q + 1], a);
                    a = fmaf(ww.z, xv[4 * q + 2], a); a = fmaf(ww.w, xv[4 * q + 3], a);
                }
                hacc[i] = a;
            }
            const uint4 u = make_uint4(th_pack_relu<F16>(hacc[0], hacc[1]), th_pack_relu<F16>(hacc[2], hacc[3]),
                                       th_pack_relu<F16>(hacc[4], hacc[5]), th_pack_relu<F16>(hacc[6], hacc[7]));
            *reinterpret_cast<uint4*>(arow + (((n8 >> 3) ^ (r & 7)) << 4)) = u;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_fmt(128, TH_OUT, F16 ? 0u : 1u);
        const uint64_t db = make_smem_desc(smem_u32(sB));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const uint64_t da = make_smem_desc(smem_u32(sA + mt * 16384));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc_mma<false>(tmem_base + mt * TH_OUT, da + 2 * kk, db + 2 * kk, idesc, kk ? 1u : 0u);
            if (F16) {
                const uint64_t dl = make_smem_desc(smem_u32(smem + TH_OFF_BLO));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tc_mma<false>(tmem_base + mt * TH_OUT, da + 2 * kk, dl + 2 * kk, idesc, 1u);
            }
        }
        tc_commit(smem_u32(bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    // ---- epilogue: thread = position row, 128 outputs -> 256 contiguous bytes ----
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * TH_OUT;
    uint16_t* orow = t1 + ((int64_t)blockIdx.x * TH_E + tid) * TH_OUT;
    uint32_t v[2][32];
    tc_ld32_issue(taddr, v[0]);
#pragma unroll
    for (int c = 0; c < TH_OUT / 32; ++c) {
        uint32_t (&cur)[32] = v[c & 1];
        tc_ld32_wait(cur);
        if (c + 1 < TH_OUT / 32) tc_ld32_issue(taddr + (c + 1) * 32, v[(c + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const float4 ba = *reinterpret_cast<const float4*>(sb3 + c * 32 + j);
            const float4 bb = *reinterpret_cast<const float4*>(sb3 + c * 32 + j + 4);
            const uint4 u = make_uint4(th_pack<F16>(__uint_as_float(cur[j + 0]) + ba.x, __uint_as_float(cur[j + 1]) + ba.y),
                                       th_pack<F16>(__uint_as_float(cur[j + 2]) + ba.z, __uint_as_float(cur[j + 3]) + ba.w),
                                       th_pack<F16>(__uint_as_float(cur[j + 4]) + bb.x, __uint_as_float(cur[j + 5]) + bb.y),
                                       th_pack<F16>(__uint_as_float(cur[j + 6]) + bb.z, __uint_as_float(cur[j + 7]) + bb.w));
            *reinterpret_cast<uint4*>(orow + c * 32 + j) = u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TH_TMEM_COLS);
    }
}

}  // namespace

int tok_head_tc(const float* x, int G, int B, int J, const float* const* gamma, const float* const* beta, const float* w0,
                const float* b0, const float* w3, const float* b3, int r0, void* t1, cudaStream_t st, int ldx, int f16) {
    EGR_CHECK(J <= TH_J && ldx >= TH_E, EGR_ERR_UNSUPPORTED, "tok_head_tc: J=%d ldx=%d", J, ldx);
    static bool attr_set = false;
    if (!attr_set) {
        EGR_CUDA_OK(cudaFuncSetAttribute(tok_head_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TH_SMEM));
        EGR_CUDA_OK(cudaFuncSetAttribute(tok_head_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TH_SMEM_F16));
        attr_set = true;
    }
    if (f16) EGR_LAUNCH(tok_head_tc_kernel<true>, G * B, 256, TH_SMEM_F16, st, x, ldx, B, J, gamma, beta, w0, b0, w3, b3, r0, (uint16_t*)t1);
    else EGR_LAUNCH(tok_head_tc_kernel<false>, G * B, 256, TH_SMEM, st, x, ldx, B, J, gamma, beta, w0, b0, w3, b3, r0, (uint16_t*)t1);
    return EGR_OK;
}

}  // namespace egr
