// ResNet stem on the tensor cores without an im2col buffer in HBM:
//
//   s2[img][oy][ox][n] = relu( sum_{ky,kx,c} W'[n][(ky*7 + kx)*3 + c] * img[c][2*oy + ky - 3][2*ox + kx - 3] + b'[n] )     n < 64
//
// (torchvision resnet18 `conv1` 7x7 stride 2 pad 3 + eval-mode `bn1` folded into W', b' + ReLU: models/backbones/resnet.py:16-21).
// One CTA = 8 consecutive output rows of one image; per row 128 output pixels = one M=128 tcgen05 tile, N = 64, K = 147 padded to 192:
//   1. the 21 input rows those output rows need are staged ONCE in shared memory as 16-bit [row][x + 3][c] (zero rows / columns =
//      padding); weights, bias, TMEM and the barrier are set up once per CTA
//   2. per output row, thread (pixel r, half h) gathers its 96 k-values - contiguous runs of 21 per ky at element offset 6r, offsets
//      known at compile time (conflict-free 16-bit LDS: consecutive pixels are 12 bytes apart) - and writes them as 16-byte pieces
//      straight into the K-major SWIZZLE_128B A tile (3 k-blocks of 64)
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
constexpr int ST_ROWS = 8;                                                      // output rows per CTA (they share 5 of their 7 input rows)
constexpr int ST_IN_ROWS = 2 * ST_ROWS + 5;                                     // 21 staged input rows
constexpr int ST_ROW = (ST_W + 6) * 3;                                          // staged row: 262 pixels x 3 channels (16-bit)
constexpr int ST_OFF_A = 0;                                                     // [3][128 rows][128 B] = 48 KB
constexpr int ST_OFF_B = ST_KB * 128 * 128;                                     // [3][64 rows][128 B]  = 24 KB
constexpr int ST_OFF_IN = ST_OFF_B + ST_KB * ST_N * 128;                        // [21][ST_ROW] 16-bit  = 33 KB
constexpr int ST_OFF_BIAS = ST_OFF_IN + ((ST_IN_ROWS * ST_ROW * 2 + 15) / 16) * 16;
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
// staged offset of im2col element k (k = (ky*7 + kx)*3 + c): ky rows down, (kx*3 + c) elements right of the window start; the
// zero-padded tail k >= 147 reads element 0 (finite; its weights are zero)
__host__ __device__ constexpr int st_off(int k) { return k < 147 ? (k / 21) * ST_ROW + (k % 21) : 0; }

// 12 pieces (8 k-values each) of one pixel's im2col row: every staged offset is a compile-time constant
template <int H>
__device__ __forceinline__ void st_build_half(const uint16_t* __restrict__ base, uint8_t* __restrict__ sA, int r, int sw) {
#pragma unroll
    for (int p = 0; p < 12; ++p) {
        uint32_t wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = H * 96 + p * 8 + 2 * e;
            const uint32_t lo = base[st_off(k)], hi = base[st_off(k + 1)];
            wv[e] = lo | (hi << 16);
        }
        const int q = H * 12 + p;
        const int kb = q >> 3, pp = q & 7;
        *reinterpret_cast<uint4*>(sA + kb * (128 * 128) + r * 128 + ((pp ^ sw) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
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
    float* sBias = reinterpret_cast<float*>(smem + ST_OFF_BIAS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ST_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int oy0 = blockIdx.x * ST_ROWS;               // this CTA: output rows [oy0, oy0 + ST_ROWS) of one image
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
    // ---- weights [64][192] 16-bit -> K-major SWIZZLE_128B tiles (one per k-block), bias ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(w16 + (int64_t)set * ST_N * STEM_K);
        for (int i = tid; i < ST_N * (STEM_K / 8); i += 256) {
            const int n = i / (STEM_K / 8), q = i - n * (STEM_K / 8);          // piece q of row n
            const int kb = q >> 3, pp = q & 7;
            *reinterpret_cast<uint4*>(sB + kb * (ST_N * 128) + n * 128 + ((pp ^ (n & 7)) << 4)) = __ldg(src + i);
        }
        if (tid < ST_N) sBias[tid] = __ldg(bias + set * ST_N + tid);
    }
    // ---- the 21 input rows of these output rows, 16-bit, [row][(x + 3)*3 + c]; out-of-image rows / columns are the zero padding ----
    {
        const float* src = img + ((int64_t)b * V + v) * 3 * ST_W * ST_W;
        const int x4 = tid & 63;                        // 4 consecutive x per thread (one float4)
        for (int pr = tid >> 6; pr < ST_IN_ROWS * 3; pr += 4) {
            const int row = pr / 3, c = pr - row * 3;
            const int iy = 2 * oy0 + row - 3;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < ST_W) val = __ldg(reinterpret_cast<const float4*>(src + ((int64_t)c * ST_W + iy) * ST_W) + x4);
            uint16_t* d = sIn + row * ST_ROW + (4 * x4 + 3) * 3 + c;
            d[0] = st_cvt<T16>(val.x); d[3] = st_cvt<T16>(val.y); d[6] = st_cvt<T16>(val.z); d[9] = st_cvt<T16>(val.w);
        }
        for (int i = tid; i < ST_IN_ROWS * 18; i += 256) {       // 3 pad pixels x 3 channels on both sides of every row
            const int row = i / 18, j = i - row * 18;
            sIn[row * ST_ROW + (j < 9 ? j : (ST_W + 3) * 3 + (j - 9))] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int r = tid & 127, h = tid >> 7;              // A tile: thread (pixel r, half h) -> k in [96h, 96h + 96): 12 pieces of 8
    const int sw = r & 7;
    const int er = (warp & 3) * 32 + (tid & 31), ch0 = (warp >> 2) * 32;      // epilogue: TMEM lane (= pixel) er, 32 channels from ch0

#pragma unroll 1
    for (int rr = 0; rr < ST_ROWS; ++rr) {
        // ---- im2col rows of output row oy0 + rr straight into the swizzled A tile ----
        const uint16_t* base = sIn + (2 * rr) * ST_ROW + 6 * r;    // the pixel's 7x7 window starts 2r pixels = 6r elements into the rows
        if (h == 0) st_build_half<0>(base, sA, r, sw);       // warp-uniform: warps 0-3 build k < 96, warps 4-7 the rest
        else st_build_half<1>(base, sA, r, sw);
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
        mbar_wait(smem_u32(bar), rr & 1);      // the MMAs have read the A tile (it may be rebuilt) and the accumulator is complete
        tc_fence_after();
        // ---- epilogue: + bias, ReLU, 64 contiguous bytes per thread ----
        uint32_t acc[32];
        tc_ld32_issue(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ch0, acc);
        tc_ld32_wait(acc);
        T16* o = out + (((int64_t)vimg * ST_WO + (oy0 + rr)) * ST_WO + er) * ST_N + ch0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = st_pack_relu<T16>(__uint_as_float(acc[j + 0]) + sBias[ch0 + j + 0], __uint_as_float(acc[j + 1]) + sBias[ch0 + j + 1]);
            u.y = st_pack_relu<T16>(__uint_as_float(acc[j + 2]) + sBias[ch0 + j + 2], __uint_as_float(acc[j + 3]) + sBias[ch0 + j + 3]);
            u.z = st_pack_relu<T16>(__uint_as_float(acc[j + 4]) + sBias[ch0 + j + 4], __uint_as_float(acc[j + 5]) + sBias[ch0 + j + 5]);
            u.w = st_pack_relu<T16>(__uint_as_float(acc[j + 6]) + sBias[ch0 + j + 6], __uint_as_float(acc[j + 7]) + sBias[ch0 + j + 7]);
            *reinterpret_cast<uint4*>(o + j) = u;
        }
        tc_fence_before();             // the accumulator has been read: the next row's MMAs (issued after the next barrier) may overwrite it
    }
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
    dim3 grid(ST_WO / ST_ROWS, V * B);
    if (dt == 2) EGR_LAUNCH(stem_tc_kernel<__half>, grid, 256, ST_SMEM, st, img, (const __half*)w16, bias, B, V, vpg, (__half*)out);
    else EGR_LAUNCH(stem_tc_kernel<__nv_bfloat16>, grid, 256, ST_SMEM, st, img, (const __nv_bfloat16*)w16, bias, B, V, vpg, (__nv_bfloat16*)out);
    return EGR_OK;
}

}  // namespace egr
