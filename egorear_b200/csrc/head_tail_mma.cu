// Heatmap-head tail, second generation: the bilinear x2 upsample ITSELF runs on the tensor cores (geometry 32x32 -> 64x64,
// C = 128, z fp16 channels-last; egoposeformer_heatmap_mvf_ex.py:108-110, :579-583):
//
//   hm[j][y][x] = sum_c W[j][c] * relu(up2_bilinear_align_corners(z)[y][x][c]) + b[j]        j < 15
//
// head_tail_tc.cu interpolates on the CUDA cores (4 LDS.128 + ~60 ALU instructions per 8 channels: 63 M warp instructions
// per launch, issue-bound at 0.2 of the HBM roofline).  Here a unit = 2 output rows = 128 pixels = one M=128 tile:
//   1. TMA loads the 3 source rows the unit can touch as a [96 px][128 ch] tile (two 64-channel SWIZZLE_128B halves):
//      read as an MN-MAJOR B operand (K = source pixel, N = channel) - no transpose anywhere
//   2. MMA-1: P = Wx * z[rows r0(y)], Q = Wx * z[rows r0(y) + 1] with Wx the x-interpolation matrix in INTEGER weights
//      (63 - rem, rem: exact in fp16, so P and Q are exact up to the fp32 accumulator), 2 x 4 tcgen05.mma (128 x 128 x 16).
//      Only two A matrices exist (the second output row of a unit starts at the same or at the next source row); they share
//      their first 64 rows: three 64-row blocks [X][X][Y] in shared memory
//   3. epilogue 1 (8 warps): relu((ay0 * P + ay1 * Q) / 63^2) in fp32 - ONE rounding, to fp16 - into the K-major A tile of
//   4. MMA-2: the 1x1 conv 128 -> 16 against W (fp16; `precise`: rows 16-31 of the B tile hold W_lo, N = 32, the two halves of
//      the accumulator are added in the epilogue), accumulator aliased onto the first columns of P
//   5. epilogue 2: + bias, staged in shared memory as [joint][128 px], ONE bulk tensor store per unit and output
// CTA c serves image group c % G (one weight set resident per CTA).
// Persistent, TWO CTAs per SM (101 KB of shared memory, 256 TMEM columns each) so that one CTA's MMA -> epilogue -> MMA
// hand-offs (~2500 cycles per unit) are covered by the other; warp-specialised (TMA producer / MMA issuer / 8 epilogue warps).
// Measured (B200, 256 images): 72-74 us against 115 us (precise) / 85 us for head_tail_tc.cu = 2.2 TB/s of algorithmic
// traffic.  What bounds it now is shared-memory bandwidth: per unit the tensor core reads 64 KB of operands for MMA-1 and
// 36 KB for MMA-2, the epilogue writes 32 + 12 KB and TMA 24 KB - ~170 KB through a 128 B/clk port = 1300 of the 2350 cycles
// a unit takes per SM (clock64 traces: shared stores of one CTA queue behind the operand reads of the other).  Next step
// would be the interpolation matrices as a TMEM-resident A operand.
// Non-finite inputs: a NaN / Inf anywhere in the 3 source rows of a unit reaches all of its 128 pixels (0 * NaN in the dense
// interpolation matrix), where the reference confines it to the pixels whose 4 taps touch it.
#include "gemm.cuh"
#include "layout_ops.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <cuda_fp16.h>
#include <type_traits>

namespace egr {
using namespace tcx;
namespace {

constexpr int M2_FS = 32, M2_FO = 64, M2_C = 128, M2_NJ = 16;
constexpr int M2_THREADS = 320, M2_EPI_WARPS = 8;
constexpr int M2_ZPX = 3 * M2_FS;                          // source pixels per unit
constexpr int M2_ZHALF = M2_ZPX * 128;                     // one 64-channel half of the z tile: 12 KB
constexpr int M2_ZSTAGE = 2 * M2_ZHALF;
constexpr int M2_OFF_PAT = 0;                              // [X][X][Y]: 3 blocks of [64 rows][128 B]
constexpr int M2_OFF_W = M2_OFF_PAT + 3 * 8192;            // [2 k-blocks][32 rows: W_hi | W_lo][128 B] of this CTA's group
constexpr int M2_OFF_A2 = M2_OFF_W + 8192;                 // [2 k-blocks][128 rows][128 B]
constexpr int M2_OFF_Z = M2_OFF_A2 + 2 * 16384;            // one stage: the other CTA of the SM covers the load latency
constexpr int M2_OFF_OUT = M2_OFF_Z + M2_ZSTAGE;           // [16 joints][128 px] fp32
constexpr int M2_OFF_OUTT = M2_OFF_OUT + M2_NJ * 512;      // [16 joints][128 px] 16-bit
constexpr int M2_OFF_BAR = M2_OFF_OUTT + M2_NJ * 256;
constexpr int M2_OFF_BIAS = M2_OFF_BAR + 64;               // [16] fp32
constexpr int M2_SMEM = M2_OFF_BIAS + M2_NJ * 4 + 1024 /*align*/;
static_assert(2 * (M2_SMEM + 1024) <= 233472, "two CTAs per SM");
constexpr uint32_t M2_TMEM_COLS = 256, M2_COL_P = 0, M2_COL_Q = 128, M2_COL_D = 0;

// MN-major SWIZZLE_128B operand: 128-byte rows = 64 elements along N, consecutive rows = consecutive K, 8-row atoms 1024 B
// apart (SBO); the next 64 elements along N start `lbo_bytes` further (LBO)
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16x2_rn(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// align_corners x2: src = dst * 31 / 63 -> integer row / column i0 and remainder rem (weights (63 - rem) / 63, rem / 63)
__device__ __forceinline__ void up_int(int dst, int& i0, int& rem) {
    const int t = dst * (M2_FS - 1);
    i0 = t / (M2_FO - 1);
    rem = t - i0 * (M2_FO - 1);
}

// Units of one group are walked by the CTAs of that group (CTA c serves group c % G: one resident weight set per CTA)
struct M2Unit { int img, b, y0, base, pat; };
__device__ __forceinline__ M2Unit m2_unit(int g, int lu, int B) {
    M2Unit r;
    r.b = lu >> 5;
    r.img = g * B + r.b;
    r.y0 = 2 * (lu & 31);
    int rem, i1;
    up_int(r.y0, r.base, rem);
    up_int(r.y0 + 1, i1, rem);
    r.pat = i1 - r.base;
    return r;
}

template <bool PRECISE>
__global__ void __launch_bounds__(M2_THREADS, 2)
head_tail_mma_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmT,
                     const float* __restrict__ w, const float* __restrict__ bias, int4 wsel, int B, int G, int J, int has_t) {
    constexpr int N2 = PRECISE ? 32 : 16;
    extern __shared__ __align__(1024) uint8_t m2_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(m2_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M2_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    float* sbias = reinterpret_cast<float*>(smem + M2_OFF_BIAS);
    const uint32_t z_full = smem_u32(bars), z_empty = z_full + 8, pq_full = z_full + 16, a2_full = z_full + 24;
    const uint32_t d2_full = z_full + 32, d2_empty = z_full + 40;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % G, ci = blockIdx.x / G;                      // this CTA's group, its rank among the group's CTAs
    const int n_cta = ((int)gridDim.x - g + G - 1) / G;                     // CTAs serving group g
    const int n_units = B * 32;                                             // units of one group
    const int sel = (g == 0) ? wsel.x : (g == 1) ? wsel.y : (g == 2) ? wsel.z : wsel.w;

    if (tid == 0) {
        mbar_init(z_full, 1); mbar_init(z_empty, 1);
        mbar_init(pq_full, 1); mbar_init(a2_full, M2_EPI_WARPS);
        mbar_init(d2_full, 1); mbar_init(d2_empty, M2_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmZ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmT)) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), M2_TMEM_COLS);

    // ---- constants (weights only: may run before the previous kernel of the stream has finished) ----
    for (int i = tid; i < 3 * 8192 / 16; i += M2_THREADS) reinterpret_cast<uint4*>(smem + M2_OFF_PAT)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < M2_NJ) sbias[tid] = (tid < J) ? __ldg(bias + (int64_t)sel * J + tid) : 0.f;
    if (tid < 256) {
        // weight tile: fp32 [J][128] -> fp16 rows 0-15 (+ the rounding residual in rows 16-31), rows >= J zero
        const int n = tid >> 4, piece = tid & 15;
        uint4 u = make_uint4(0u, 0u, 0u, 0u), ul = make_uint4(0u, 0u, 0u, 0u);
        if (n < J) {
            const float4* src = reinterpret_cast<const float4*>(w + ((int64_t)sel * J + n) * M2_C + piece * 8);
            const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
            u = make_uint4(pack_f16x2_rn(f0.x, f0.y), pack_f16x2_rn(f0.z, f0.w), pack_f16x2_rn(f1.x, f1.y), pack_f16x2_rn(f1.z, f1.w));
            if (PRECISE) {
                const __half2* h = reinterpret_cast<const __half2*>(&u);
                const float2 a = __half22float2(h[0]), b2 = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
                ul = make_uint4(pack_f16x2_rn(f0.x - a.x, f0.y - a.y), pack_f16x2_rn(f0.z - b2.x, f0.w - b2.y),
                                pack_f16x2_rn(f1.x - c.x, f1.y - c.y), pack_f16x2_rn(f1.z - d.x, f1.w - d.y));
            }
        }
        const int kb = piece >> 3, pp = piece & 7;
        uint8_t* wt = smem + M2_OFF_W + kb * 4096;
        *reinterpret_cast<uint4*>(wt + n * 128 + ((pp ^ (n & 7)) << 4)) = u;
        *reinterpret_cast<uint4*>(wt + (16 + n) * 128 + ((pp ^ (n & 7)) << 4)) = ul;
    }
    __syncthreads();
    if (tid < 192) {
        // x-interpolation blocks, K-major [64 rows = x][K = 2 source rows x 32 px]: X (blocks 0, 1) reads source row 0 of the
        // window, Y (block 2) source row 1.  A-matrix of pattern p = the 128 rows starting at block p: output row 0 always
        // on source row 0, output row 1 on source row p
        const int blk = tid >> 6, x = tid & 63;
        int i0, rem;
        up_int(x, i0, rem);
        const int k0 = (blk == 2 ? M2_FS : 0) + i0;
        __half* row = reinterpret_cast<__half*>(smem + M2_OFF_PAT + blk * 8192 + x * 128);
        row[((((k0 >> 3) ^ (x & 7)) << 3)) + (k0 & 7)] = __float2half_rn((float)(M2_FO - 1 - rem));
        if (rem) {
            const int k1 = k0 + 1;
            row[((((k1 >> 3) ^ (x & 7)) << 3)) + (k1 & 7)] = __float2half_rn((float)rem);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // after the TMEM allocation (common.cuh)
    pdl_wait();

    if (warp == 0) {
        // ================= TMA producer: the 3-row z window of every unit (the whole warp runs the loop, one lane issues) =======
        int it = 0;
        for (int lu = ci; lu < n_units; lu += n_cta, ++it) {
            const M2Unit un = m2_unit(g, lu, B);
            mbar_wait(z_empty, (it & 1) ^ 1);
            const int px0 = un.img * (M2_FS * M2_FS) + un.base * M2_FS;      // rows past the tensor end are zero-filled
            if (elect_one()) {
                mbar_expect_tx(z_full, M2_ZSTAGE);
                tma_load_3d(sbase + M2_OFF_Z, &tmZ, z_full, 0, px0, 0);
                tma_load_3d(sbase + M2_OFF_Z + M2_ZHALF, &tmZ, z_full, 64, px0, 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================= MMA issuer (the whole warp runs the loop, one lane issues) =================
        constexpr uint32_t idesc1 = make_idesc_fmt(128, M2_C, 0u) | (1u << 16);      // B operand MN-major
        constexpr uint32_t idesc2 = make_idesc_fmt(128, N2, 0u);
        int it = 0;
        for (int lu = ci; lu < n_units; lu += n_cta, ++it) {
            const M2Unit un = m2_unit(g, lu, B);
            const uint32_t par = it & 1;
            mbar_wait(z_full, par);
            mbar_wait(d2_empty, par ^ 1);      // the previous unit's D (aliased onto P) has been read; its P / Q long before
            tc_fence_after();
            const uint64_t da = make_smem_desc(sbase + M2_OFF_PAT + un.pat * 8192);
            const uint64_t dz = make_smem_desc_mn(sbase + M2_OFF_Z, M2_ZHALF);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)      // 16 source pixels per MMA = 2 atoms of 8 K-rows = 2048 B (encoded >> 4)
                    tc_mma<false>(tmem_base + M2_COL_P, da + 2 * kk, dz + 128 * kk, idesc1, kk ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)      // the same window one source row (32 px = 4 KB) further
                    tc_mma<false>(tmem_base + M2_COL_Q, da + 2 * kk, dz + 256 + 128 * kk, idesc1, kk ? 1u : 0u);
                tc_commit(z_empty);
                tc_commit(pq_full);
            }
            __syncwarp();
            mbar_wait(a2_full, par);           // epilogue 1 has drained P / Q and written the A tile
            tc_fence_after();
            const uint64_t d2a = make_smem_desc(sbase + M2_OFF_A2), d2b = make_smem_desc(sbase + M2_OFF_W);
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)      // k-blocks: 16 KB apart in the A tile, 4 KB apart in the weight tile
                        tc_mma<false>(tmem_base + M2_COL_D, d2a + 1024 * kb + 2 * kk, d2b + 256 * kb + 2 * kk, idesc2, (kb | kk) ? 1u : 0u);
                tc_commit(d2_full);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue warps =================
        const int e = warp - 2;
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        const int half = e >> 2;           // channel half (epilogue 1) / joint half (epilogue 2)
        const int r = q * 32 + lane, yy = r >> 6;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t a2row = sbase + M2_OFF_A2 + half * 16384 + r * 128;
        const float inv = 1.f / (float)((M2_FO - 1) * (M2_FO - 1));
        const uint32_t so32 = sbase + M2_OFF_OUT + r * 4, st32 = sbase + M2_OFF_OUTT + r * 2;      // staging [joint][128 px]
        const bool issuer = (e == 0) && (lane == 0);
        int it = 0;
        for (int lu = ci; lu < n_units; lu += n_cta, ++it) {
            const M2Unit un = m2_unit(g, lu, B);
            const uint32_t par = it & 1;
            int iy, remy;
            up_int(un.y0 + yy, iy, remy);
            const float fy0 = (float)(M2_FO - 1 - remy) * inv, fy1 = (float)remy * inv;
            mbar_wait(pq_full, par);           // (the A tile is free: this thread waited for the previous unit's d2_full)
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t vp[16], vq[16];
                tc_ld16_issue(lane_addr + M2_COL_P + half * 64 + c * 16, vp);
                tc_ld16_issue(lane_addr + M2_COL_Q + half * 64 + c * 16, vq);
                tc_ld16_wait(vp);
                tc_ld16_wait(vq);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = j * 8 + k * 2;
                        float a = fy0 * __uint_as_float(vp[i]), b2 = fy0 * __uint_as_float(vp[i + 1]);
                        if (remy) {       // weight 0 rows are skipped, not multiplied (the window may end past the image)
                            a = fmaf(fy1, __uint_as_float(vq[i]), a);
                            b2 = fmaf(fy1, __uint_as_float(vq[i + 1]), b2);
                        }
                        o[k] = pack_relu_f16x2(a, b2);
                    }
                    const int pp = c * 2 + j;
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a2row + ((pp ^ (r & 7)) << 4)), "r"(o[0]), "r"(o[1]),
                                 "r"(o[2]), "r"(o[3]) : "memory");
                }
            }
            tc_fence_before();
            fence_async_smem();            // generic-proxy writes of the A tile -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(a2_full);

            mbar_wait(d2_full, par);
            tc_fence_after();
            uint32_t v[8], vl[8];
            tc_ld8_issue(lane_addr + M2_COL_D + half * 8, v);
            if (PRECISE) tc_ld8_issue(lane_addr + M2_COL_D + 16 + half * 8, vl);
            tc_ld8_wait(v);
            if (PRECISE) tc_ld8_wait(vl);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty);
            // staging tiles [joint][128 px]: free once the previous unit's bulk stores have read them
            if (issuer) tma_store_wait_read();
            named_bar_sync_1<M2_EPI_WARPS * 32>();
            // explicit shared-space accesses: through generic pointers every bias load would wait for the store before it
            float bq[8];
            {
                const uint4 b0 = lds128(sbase + M2_OFF_BIAS + (half * 8) * 4);
                const uint4 b1 = lds128(sbase + M2_OFF_BIAS + (half * 8 + 4) * 4);
                bq[0] = __uint_as_float(b0.x); bq[1] = __uint_as_float(b0.y); bq[2] = __uint_as_float(b0.z); bq[3] = __uint_as_float(b0.w);
                bq[4] = __uint_as_float(b1.x); bq[5] = __uint_as_float(b1.y); bq[6] = __uint_as_float(b1.z); bq[7] = __uint_as_float(b1.w);
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float val = __uint_as_float(v[jj]) + bq[jj];
                if (PRECISE) val += __uint_as_float(vl[jj]);
                unsigned short h16;
                if (PRECISE) asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h16) : "f"(val));
                else h16 = __bfloat16_as_ushort(__float2bfloat16_rn(val));
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(so32 + (half * 8 + jj) * 512), "f"(val) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(st32 + (half * 8 + jj) * 256), "h"(h16) : "memory");
            }
            fence_async_smem();
            named_bar_sync_1<M2_EPI_WARPS * 32>();
            if (issuer) {
                tma_store_4d(&tmO, sbase + M2_OFF_OUT, un.y0 * M2_FO, 0, g, un.b);
                if (has_t) tma_store_3d(&tmT, sbase + M2_OFF_OUTT, un.y0 * M2_FO, 0, un.img);
            }
        }
        if (issuer) tma_store_wait_all();      // the staging tiles must outlive the last bulk stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, M2_TMEM_COLS);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// R1 tail on the same machinery: refn = relu(up2_bilinear_align_corners(z)) as a 16-bit channels-last map (the input of the
// H2 3x3 conv, and in EGR_PREC_FP16 the copy pose3d samples).  MMA-1 and the fp32 blend are those of the head tail; the
// K-major SWIZZLE_128B tile the blend writes is exactly the layout a bulk tensor store takes, so the unit leaves as two
// [128 px][64 ch] boxes - no second MMA, no staging pass.
constexpr int U2_OFF_PAT = 0;
constexpr int U2_OFF_A2 = U2_OFF_PAT + 3 * 8192;           // [2 channel halves][128 rows][128 B]
constexpr int U2_OFF_Z = U2_OFF_A2 + 2 * 16384;            // two stages
constexpr int U2_STAGES = 2;
constexpr int U2_OFF_BAR = U2_OFF_Z + U2_STAGES * M2_ZSTAGE;
constexpr int U2_SMEM = U2_OFF_BAR + 64 + 1024 /*align*/;
static_assert(2 * (U2_SMEM + 1024) <= 233472, "two CTAs per SM");

template <bool OUT_F16>
__global__ void __launch_bounds__(M2_THREADS, 2)
up2_relu_mma_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmO, int n_img) {
    extern __shared__ __align__(1024) uint8_t m2_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(m2_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + U2_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
    const uint32_t z_full0 = smem_u32(bars), z_empty0 = z_full0 + 16, pq_full = z_full0 + 32, pq_drained = z_full0 + 40;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_units = n_img * 32;

    if (tid == 0) {
        for (int s2 = 0; s2 < U2_STAGES; ++s2) { mbar_init(z_full0 + 8 * s2, 1); mbar_init(z_empty0 + 8 * s2, 1); }
        mbar_init(pq_full, 1); mbar_init(pq_drained, M2_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmZ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), M2_TMEM_COLS);
    for (int i = tid; i < 3 * 8192 / 16; i += M2_THREADS) reinterpret_cast<uint4*>(smem + U2_OFF_PAT)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid < 192) {      // the x-interpolation blocks [X][X][Y] of head_tail_mma_kernel
        const int blk = tid >> 6, x = tid & 63;
        int i0, rem;
        up_int(x, i0, rem);
        const int k0 = (blk == 2 ? M2_FS : 0) + i0;
        __half* row = reinterpret_cast<__half*>(smem + U2_OFF_PAT + blk * 8192 + x * 128);
        row[((((k0 >> 3) ^ (x & 7)) << 3)) + (k0 & 7)] = __float2half_rn((float)(M2_FO - 1 - rem));
        if (rem) {
            const int k1 = k0 + 1;
            row[((((k1 >> 3) ^ (x & 7)) << 3)) + (k1 & 7)] = __float2half_rn((float)rem);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        int stage = 0, phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const M2Unit un = m2_unit(0, u, n_img);
            mbar_wait(z_empty0 + 8 * stage, phase ^ 1);
            const uint32_t full = z_full0 + 8 * stage;
            const uint32_t dst = sbase + U2_OFF_Z + stage * M2_ZSTAGE;
            const int px0 = un.img * (M2_FS * M2_FS) + un.base * M2_FS;
            if (elect_one()) {
                mbar_expect_tx(full, M2_ZSTAGE);
                tma_load_3d(dst, &tmZ, full, 0, px0, 0);
                tma_load_3d(dst + M2_ZHALF, &tmZ, full, 64, px0, 0);
            }
            __syncwarp();
            if (++stage == U2_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc1 = make_idesc_fmt(128, M2_C, 0u) | (1u << 16);      // B operand MN-major
        int stage = 0, phase = 0, it = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
            const M2Unit un = m2_unit(0, u, n_img);
            mbar_wait(z_full0 + 8 * stage, phase);
            mbar_wait(pq_drained, (it & 1) ^ 1);      // the epilogue has read the previous unit's P / Q
            tc_fence_after();
            const uint64_t da = make_smem_desc(sbase + U2_OFF_PAT + un.pat * 8192);
            const uint64_t dz = make_smem_desc_mn(sbase + U2_OFF_Z + stage * M2_ZSTAGE, M2_ZHALF);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tc_mma<false>(tmem_base + M2_COL_P, da + 2 * kk, dz + 128 * kk, idesc1, kk ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tc_mma<false>(tmem_base + M2_COL_Q, da + 2 * kk, dz + 256 + 128 * kk, idesc1, kk ? 1u : 0u);
                tc_commit(z_empty0 + 8 * stage);
                tc_commit(pq_full);
            }
            __syncwarp();
            if (++stage == U2_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        const int e = warp - 2, q = warp & 3, half = e >> 2;
        const int r = q * 32 + lane, yy = r >> 6;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t a2row = sbase + U2_OFF_A2 + half * 16384 + r * 128;
        const float inv = 1.f / (float)((M2_FO - 1) * (M2_FO - 1));
        const bool issuer = (e == 0) && (lane == 0);
        int it = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
            const M2Unit un = m2_unit(0, u, n_img);
            int iy, remy;
            up_int(un.y0 + yy, iy, remy);
            const float fy0 = (float)(M2_FO - 1 - remy) * inv, fy1 = (float)remy * inv;
            mbar_wait(pq_full, it & 1);
            tc_fence_after();
            // the tile is the source of the previous unit's bulk stores: wait until they have read it
            if (issuer) tma_store_wait_read();
            named_bar_sync_1<M2_EPI_WARPS * 32>();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t vp[16], vq[16];
                tc_ld16_issue(lane_addr + M2_COL_P + half * 64 + c * 16, vp);
                tc_ld16_issue(lane_addr + M2_COL_Q + half * 64 + c * 16, vq);
                tc_ld16_wait(vp);
                tc_ld16_wait(vq);
                if (c == 3) {      // every accumulator column of this warp is in registers: the next unit's MMA may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pq_drained);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = j * 8 + k * 2;
                        float a = fy0 * __uint_as_float(vp[i]), b2 = fy0 * __uint_as_float(vp[i + 1]);
                        if (remy) {
                            a = fmaf(fy1, __uint_as_float(vq[i]), a);
                            b2 = fmaf(fy1, __uint_as_float(vq[i + 1]), b2);
                        }
                        if (OUT_F16) {
                            o[k] = pack_relu_f16x2(a, b2);
                        } else {
                            const __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(a, 0.f), fmaxf(b2, 0.f));
                            o[k] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                    }
                    const int pp = c * 2 + j;
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a2row + ((pp ^ (r & 7)) << 4)), "r"(o[0]), "r"(o[1]),
                                 "r"(o[2]), "r"(o[3]) : "memory");
                }
            }
            fence_async_smem();
            named_bar_sync_1<M2_EPI_WARPS * 32>();
            if (issuer) {
                const int px = un.img * (M2_FO * M2_FO) + un.y0 * M2_FO;      // rows y0, y0 + 1 = 128 consecutive pixels
                tma_store_3d(&tmO, sbase + U2_OFF_A2, 0, px, 0);
                tma_store_3d(&tmO, sbase + U2_OFF_A2 + 16384, 64, px, 0);
            }
        }
        if (issuer) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, M2_TMEM_COLS);
    }
}

}  // namespace

int head_tail_mma(const void* z, const float* w, const float* bias, const int* wsel_host, int B, int G, int J, float* hm,
                  int64_t hm_bs, int64_t hm_gs, void* hm_t, int precise, cudaStream_t st) {
    EGR_CHECK(J <= M2_NJ && G <= 4 && B > 0, EGR_ERR_UNSUPPORTED, "head_tail_mma: J=%d G=%d B=%d", J, G, B);
    EGR_CHECK((int64_t)G * B * 32 < (1ll << 30), EGR_ERR_UNSUPPORTED, "head_tail_mma: %d images", G * B);
    if (hm_bs % 4 || hm_gs % 4 || (uintptr_t)hm % 16 || (uintptr_t)hm_t % 16)      // bulk tensor stores need 16-byte rows
        return head_tail_tc(z, w, bias, wsel_host, B, G, J, hm, hm_bs, hm_gs, hm_t, precise, st);
    static bool attr_set = false;
    if (!attr_set) {
        EGR_CUDA_OK(cudaFuncSetAttribute(head_tail_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM));
        EGR_CUDA_OK(cudaFuncSetAttribute(head_tail_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM));
        attr_set = true;
    }
    CUtensorMap tmZ, tmO, tmT;
    int rc;
    {   // z as [1][total px][128 ch] fp16; box = 64 channels x 96 pixels (3 source rows), SWIZZLE_128B
        const uint64_t dims[3] = {(uint64_t)M2_C, (uint64_t)G * B * M2_FS * M2_FS, 1};
        const uint64_t str[2] = {(uint64_t)M2_C * 2, (uint64_t)G * B * M2_FS * M2_FS * M2_C * 2};
        const uint32_t box[3] = {64, (uint32_t)M2_ZPX, 1};
        if ((rc = tc_encode_tiled(&tmZ, DT_F16, z, 3, dims, str, box, "head tail z", false))) return rc;
    }
    {   // hm (b, g, j, px) fp32; box = 128 px x J joints, rows of 512 B (no swizzle)
        const uint64_t dims[4] = {(uint64_t)M2_FO * M2_FO, (uint64_t)J, (uint64_t)G, (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)M2_FO * M2_FO * 4, (uint64_t)hm_gs * 4, (uint64_t)hm_bs * 4};
        const uint32_t box[4] = {128, (uint32_t)J, 1, 1};
        if ((rc = tc_encode_tiled(&tmO, DT_F32, hm, 4, dims, str, box, "head tail hm", true))) return rc;
    }
    {   // hm_t [img][J][px] 16-bit (the fp32 map stands in when there is no copy: never stored through)
        const uint64_t dims[3] = {(uint64_t)M2_FO * M2_FO, (uint64_t)J, (uint64_t)G * B};
        const uint64_t str[2] = {(uint64_t)M2_FO * M2_FO * 2, (uint64_t)J * M2_FO * M2_FO * 2};
        const uint32_t box[3] = {128, (uint32_t)J, 1};
        if (hm_t) {
            if ((rc = tc_encode_tiled(&tmT, precise ? DT_F16 : DT_BF16, hm_t, 3, dims, str, box, "head tail hm_t", true))) return rc;
        } else {
            tmT = tmO;
        }
    }
    const int grid = std::min(2 * sm_count(), G * B * 32);
    const int4 sel = make_int4(wsel_host[0], G > 1 ? wsel_host[1] : 0, G > 2 ? wsel_host[2] : 0, G > 3 ? wsel_host[3] : 0);
    if (precise)
        EGR_LAUNCH(head_tail_mma_kernel<true>, grid, M2_THREADS, M2_SMEM, st, tmZ, tmO, tmT, w, bias, sel, B, G, J, hm_t ? 1 : 0);
    else
        EGR_LAUNCH(head_tail_mma_kernel<false>, grid, M2_THREADS, M2_SMEM, st, tmZ, tmO, tmT, w, bias, sel, B, G, J, hm_t ? 1 : 0);
    return EGR_OK;
}


// relu(up2(z)) -> one 16-bit channels-last map (R1 tail when no other output is wanted); returns EGR_ERR_UNSUPPORTED-free:
// callers check the shape / alignment conditions themselves (layout_ops.cu up2_relu_dual)
int up2_relu_mma(const void* z, int n_img, void* out, int out_f16, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        EGR_CUDA_OK(cudaFuncSetAttribute(up2_relu_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_SMEM));
        EGR_CUDA_OK(cudaFuncSetAttribute(up2_relu_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_SMEM));
        attr_set = true;
    }
    CUtensorMap tmZ, tmO;
    int rc;
    {
        const uint64_t dims[3] = {(uint64_t)M2_C, (uint64_t)n_img * M2_FS * M2_FS, 1};
        const uint64_t str[2] = {(uint64_t)M2_C * 2, (uint64_t)n_img * M2_FS * M2_FS * M2_C * 2};
        const uint32_t box[3] = {64, (uint32_t)M2_ZPX, 1};
        if ((rc = tc_encode_tiled(&tmZ, DT_F16, z, 3, dims, str, box, "up2 z", false))) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)M2_C, (uint64_t)n_img * M2_FO * M2_FO, 1};
        const uint64_t str[2] = {(uint64_t)M2_C * 2, (uint64_t)n_img * M2_FO * M2_FO * M2_C * 2};
        const uint32_t box[3] = {64, 128, 1};
        if ((rc = tc_encode_tiled(&tmO, out_f16 ? DT_F16 : DT_BF16, out, 3, dims, str, box, "up2 out", false))) return rc;
    }
    const int grid = std::min(2 * sm_count(), n_img * 32);
    if (out_f16) EGR_LAUNCH(up2_relu_mma_kernel<true>, grid, M2_THREADS, U2_SMEM, st, tmZ, tmO, n_img);
    else EGR_LAUNCH(up2_relu_mma_kernel<false>, grid, M2_THREADS, U2_SMEM, st, tmZ, tmO, n_img);
    return EGR_OK;
}

}  // namespace egr
