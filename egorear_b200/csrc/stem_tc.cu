// ResNet stem on the tensor cores without an im2col buffer in HBM:
//
//   s2[img][oy][ox][n] = relu( sum_{ky,kx,c} W'[n][(ky*7 + kx)*3 + c] * img[c][2*oy + ky - 3][2*ox + kx - 3] + b'[n] )     n < 64
//
// (torchvision resnet18 `conv1` 7x7 stride 2 pad 3 + eval-mode `bn1` folded into W', b' + ReLU: models/backbones/resnet.py:16-21).
// One CTA = one output row of one image = 128 output pixels = one M=128 tcgen05 tile, N = 64, K = 147 padded to 192:
//   1. the 7 input rows the output row needs are staged in shared memory as 16-bit [ky][x + 3][c] (zero rows / columns = padding)
//   2. thread (pixel r, half h) gathers its 96 k-values - contiguous runs of 21 per ky at element offset 6r - through a k -> offset
//      table (conflict-free 16-bit LDS: consecutive pixels are 12 bytes apart) and writes them as 16-byte pieces straight into the
//      K-major SWIZZLE_128B A tile (3 k-blocks of 64)
//   3. one thread issues 3 x 4 tcgen05.mma (M=128, N=64, K=16) against the resident weight tile; accumulator in TMEM
//   4. tcgen05.ld gives thread (pixel r, column half) 32 channels: + bias, ReLU, 64 contiguous bytes per thread
// The im2col path it replaces wrote and re-read 6.3 MB per image; this reads the image once (0.75 MB) and writes the map (2 MB).
#include "backbone_ops.cuh"
#include "tc_ptx.cuh"
#include <type_traits>

namespace egr {
using namespace tcx;
namespace {

constexpr int ST_W = 256, ST_WO = 128, ST_N = 64, ST_KB = STEM_K / 64;          // 3 k-blocks
constexpr int ST_ROW = (ST_W + 6) * 3;                                          // staged row: 262 pixels x 3 channels (16-bit)
constexpr int ST_OFF_A = 0;                                                     // [3][128 rows][128 B] = 48 KB
constexpr int ST_OFF_B = ST_KB * 128 * 128;                                     // [3][64 rows][128 B]  = 24 KB
constexpr int ST_OFF_IN = ST_OFF_B + ST_KB * ST_N * 128;                        // [7][ST_ROW] 16-bit   = 11 KB
constexpr int ST_OFF_TAB = ST_OFF_IN + ((7 * ST_ROW * 2 + 15) / 16) * 16;       // [192] int16: k -> ky*ST_ROW + (k % 21), -1 for the zero pad
constexpr int ST_OFF_BIAS = ST_OFF_TAB + STEM_K * 2;
constexpr int ST_OFF_BAR = ST_OFF_BIAS + ST_N * 4;
constexpr int ST_SMEM = ST_OFF_BAR + 64 + 1024 /*align*/;
constexpr uint32_t ST_TMEM_COLS = 64;

template <typename T16> __device__ __forceinline__ uint16_t st_cvt(float x);
template <> __device__ __forceinline__ uint16_t st_cvt<__nv_bfloat16>(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
template <> __device__ __forceinline__ uint16_t st_cvt<__half>(float x) {
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    return h;
}
template <typename T16> __device__ __forceinline__ uint32_t st_pack_relu(float lo, float hi);
template <> __device__ __forceinline__ uint32_t st_pack_relu<__nv_bfloat16>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <> __device__ __forceinline__ uint32_t st_pack_relu<__half>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <typename T16>
__global__ void __launch_bounds__(256, 2)
stem_tc_kernel(const float* __restrict__ img, const T16* __restrict__ w16, const float* __restrict__ bias, int B, int V, int vpg,
               T16* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t st_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem + ST_OFF_A;
    uint8_t* sB = smem + ST_OFF_B;
    uint16_t* sIn = reinterpret_cast<uint16_t*>(smem + ST_OFF_IN);
    int16_t* sTab = reinterpret_cast<int16_t*>(smem + ST_OFF_TAB);
    float* sBias = reinterpret_cast<float*>(smem + ST_OFF_BIAS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ST_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int oy = blockIdx.x;
    const int vimg = blockIdx.y;                        // view-major image index v * B + b (the engines' activation order)
    const int v = vimg / B, b = vimg - v * B;
    const int set = v / vpg;                            // front weights: views [0, vpg), back weights: the rest

    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), ST_TMEM_COLS);
        pdl_trigger();          // only after this CTA owns its TMEM columns (common.cuh)
    }
    pdl_wait();
    // ---- weights [64][192] 16-bit -> K-major SWIZZLE_128B tiles (one per k-block), bias, k -> offset table ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(w16 + (int64_t)set * ST_N * STEM_K);
        for (int i = tid; i < ST_N * (STEM_K / 8); i += 256) {
            const int n = i / (STEM_K / 8), q = i - n * (STEM_K / 8);          // piece q of row n
            const int kb = q >> 3, pp = q & 7;
            *reinterpret_cast<uint4*>(sB + kb * (ST_N * 128) + n * 128 + ((pp ^ (n & 7)) << 4)) = __ldg(src + i);
        }
        if (tid < ST_N) sBias[tid] = __ldg(bias + set * ST_N + tid);
        if (tid < STEM_K) {
            const int ky = tid / 21, rem = tid - ky * 21;
            sTab[tid] = (tid < 147) ? (int16_t)(ky * ST_ROW + rem) : (int16_t)-1;
        }
    }
    // ---- the 7 input rows of this output row, 16-bit, [ky][(x + 3)*3 + c]; out-of-image rows / columns are the zero padding ----
    {
        const float* src = img + ((int64_t)b * V + v) * 3 * ST_W * ST_W;
        for (int i = tid; i < 7 * 3 * ST_W; i += 256) {
            const int x = i & (ST_W - 1);
            const int c = (i >> 8) % 3, ky = i / (3 * ST_W);
            const int iy = 2 * oy + ky - 3;
            const float val = (iy >= 0 && iy < ST_W) ? __ldg(src + ((int64_t)c * ST_W + iy) * ST_W + x) : 0.f;
            sIn[ky * ST_ROW + (x + 3) * 3 + c] = st_cvt<T16>(val);
        }
        for (int i = tid; i < 7 * 18; i += 256) {       // 3 pad pixels x 3 channels on both sides of every row
            const int ky = i / 18, j = i - ky * 18;
            sIn[ky * ST_ROW + (j < 9 ? j : (ST_W + 3) * 3 + (j - 9))] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- A tile: thread (pixel r, half h) -> k in [96h, 96h + 96): 12 pieces of 8 ----
    {
        const int r = tid & 127, h = tid >> 7;
        const uint16_t* base = sIn + 6 * r;             // pixel r's window starts 2r pixels = 6r elements into every staged row
        const int sw = r & 7;
#pragma unroll 4
        for (int p = 0; p < 12; ++p) {
            const int q = h * 12 + p, k0 = q * 8;
            uint32_t wv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int o0 = sTab[k0 + 2 * e], o1 = sTab[k0 + 2 * e + 1];
                const uint32_t lo = (o0 >= 0) ? base[o0] : 0u, hi = (o1 >= 0) ? base[o1] : 0u;
                wv[e] = lo | (hi << 16);
            }
            const int kb = q >> 3, pp = q & 7;
            *reinterpret_cast<uint4*>(sA + kb * (128 * 128) + r * 128 + ((pp ^ sw) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
    fence_async_smem();            // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_fmt(128, ST_N, std::is_same<T16, __half>::value ? 0u : 1u);
#pragma unroll
        for (int kb = 0; kb < ST_KB; ++kb) {
            const uint64_t da = make_smem_desc(smem_u32(sA + kb * (128 * 128)));
            const uint64_t db = make_smem_desc(smem_u32(sB + kb * (ST_N * 128)));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc_mma<false>(tmem_base, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
        }
        tc_commit(smem_u32(bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    // ---- epilogue: thread = (pixel row of its TMEM lane quarter, 32-channel half) ----
    {
        const int r = (warp & 3) * 32 + (tid & 31), ch0 = (warp >> 2) * 32;
        uint32_t acc[32];
        tc_ld32_issue(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ch0, acc);
        tc_ld32_wait(acc);
        T16* o = out + (((int64_t)vimg * ST_WO + oy) * ST_WO + r) * ST_N + ch0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = st_pack_relu<T16>(__uint_as_float(acc[j + 0]) + sBias[ch0 + j + 0], __uint_as_float(acc[j + 1]) + sBias[ch0 + j + 1]);
            u.y = st_pack_relu<T16>(__uint_as_float(acc[j + 2]) + sBias[ch0 + j + 2], __uint_as_float(acc[j + 3]) + sBias[ch0 + j + 3]);
            u.z = st_pack_relu<T16>(__uint_as_float(acc[j + 4]) + sBias[ch0 + j + 4], __uint_as_float(acc[j + 5]) + sBias[ch0 + j + 5]);
            u.w = st_pack_relu<T16>(__uint_as_float(acc[j + 6]) + sBias[ch0 + j + 6], __uint_as_float(acc[j + 7]) + sBias[ch0 + j + 7]);
            *reinterpret_cast<uint4*>(o + j) = u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ST_TMEM_COLS);
    }
}

}  // namespace

int stem_tc(const float* img, const void* w16, const float* bias, int dt, int B, int V, int vpg, void* out, cudaStream_t st) {
    EGR_CHECK(dt == 1 || dt == 2, EGR_ERR_UNSUPPORTED, "stem_tc: dt=%d", dt);
    EGR_CHECK((int64_t)V * B <= 65535, EGR_ERR_UNSUPPORTED, "stem_tc: %lld images exceed gridDim.y; split the batch", (long long)V * B);
    static bool attr_set = false;
    if (!attr_set) {
        EGR_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
        EGR_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
        attr_set = true;
    }
    dim3 grid(ST_WO, V * B);
    if (dt == 2) EGR_LAUNCH(stem_tc_kernel<__half>, grid, 256, ST_SMEM, st, img, (const __half*)w16, bias, B, V, vpg, (__half*)out);
    else EGR_LAUNCH(stem_tc_kernel<__nv_bfloat16>, grid, 256, ST_SMEM, st, img, (const __nv_bfloat16*)w16, bias, B, V, vpg, (__nv_bfloat16*)out);
    return EGR_OK;
}

}  // namespace egr
