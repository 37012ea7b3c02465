// tcgen05 + TMA GEMM / implicit-GEMM conv: the EGR_PREC_BF16 dense stage (sm_100a only).
//
//   D[M,N] = epi(A[M,K] · W[N,K]^T + bias)      A, W bf16 (kind::f16) or fp32 read as TF32 (kind::tf32),
//                                               K contiguous, fp32 accumulate in TMEM
//
// One persistent CTA per SM walks a static tile list (n fastest, so the CTAs that share an A tile run together
// and hit L2).  Roles (warp-specialised, mbarrier pipelines, no __syncthreads in the main loop):
//   warp 0      TMA producer: A tile 128x64 and W tile BNx64 per k-block into a SWIZZLE_128B smem ring
//   warp 1      TMEM allocator + MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block,
//               tcgen05.commit releases the smem slot / publishes the accumulator
//   warps 2-9   epilogue: tcgen05.ld the 128xBN fp32 accumulator (two TMEM buffers, so the epilogue of tile i
//               overlaps the MMAs of tile i+1; the load of the next 32-column chunk is in flight while one is
//               processed), bias from smem + activation (+ fused bilinear-x2 residual); rows are staged in a
//               swizzled smem tile and leave through TMA stores (whole 128 B row segments, rows >= M clipped)
// The 3x3 stride-2 pad-1 conv is an implicit GEMM: the NHWC input is described to TMA as a 5-D tensor
// (2*Cin, W/2, 2, H/2, img) so that tap (ky,kx) of 128 consecutive output pixels is ONE box load whose
// out-of-bounds part (the zero padding) is filled by the TMA unit.
#include "gemm.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <mutex>
#include <unordered_map>
#include <type_traits>

namespace egr {
using namespace tcx;
int g_opt_ws = 0;      // weight-stationary mode of gemm_tc (option "ws"): measured no gain on B200, off by default
// option "conv_prefetch": L2 prefetch (cp.async.bulk.prefetch.tensor) of the next tile's input rows in the 3x3 stride-1 conv.
// Measured on the backbone at 512 frames: layer1 7.18 -> 8.10 ms, FPN 7.34 -> 7.63 ms: the convs are not waiting on DRAM
// misses (what bounds them is the 9x re-read of the input through L2 -> shared memory), so it is off by default.
int g_opt_conv_prefetch = 0;
int g_opt_pair = 0;            // option "pair": split weights as [A | W_hi | W_lo] ring stages (BN <= 128) instead of the K wrap
                               // (-0.4 % step time; off by default: every parity figure in DESIGN.md was taken with the K wrap)
namespace {

constexpr int BM = 128;
constexpr int ROW_BYTES = 128;                       // one SWIZZLE_128B row = one k-block: 64 bf16 or 32 fp32
constexpr int EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + EPI_WARPS * 32;
constexpr int SMEM_RING = 192 * 1024;

template <int BN> struct TcCfg {
    static constexpr int STAGE_A = BM * ROW_BYTES;
    static constexpr int STAGE_B = BN * ROW_BYTES;
    static constexpr int STAGE = STAGE_A + STAGE_B;
    static constexpr int STAGES = SMEM_RING / STAGE;          // 256: 4, 128: 6, 64: 8
    static constexpr int TMEM_COLS = 2 * BN;                  // double-buffered accumulator
    // ring | store staging (8 warps x 4 KB, 1024-aligned: TMA-store source) | barriers | bias tile
    static constexpr int OFF_STAGING = STAGES * STAGE;
    static constexpr int OFF_BARS = OFF_STAGING + EPI_WARPS * 4096;
    static constexpr int OFF_BIAS = OFF_BARS + 256;
    static constexpr int SMEM = OFF_BIAS + 1024 /*bias[BN <= 256]*/ + 1024 /*align*/;
};
static_assert(TcCfg<256>::SMEM <= 227 * 1024 && TcCfg<128>::SMEM <= 227 * 1024 && TcCfg<64>::SMEM <= 227 * 1024, "smem budget");

struct TcParams {
    int M, N, K;
    int m_tiles, n_tiles, groups, ksplit, kb_total, kb_per_split;
    int amode;
    int kblk;                 // A_PLAIN: elements per outer k block (== K when there is no K split of A)
    int ka;                   // A's own K: k >= ka re-reads A at k - ka (split weights [W_hi | W_lo]); == K otherwise
    int Cin, Wout, HWout;     // A_CONV3S2 / A_CONV3S1 (output geometry; == input geometry for the stride-1 conv)
    int tap_fixed;            // A_CONV3S2 with K == Cin: the centre tap only (1x1 stride-2 conv); -1 otherwise
    int pf_rows;              // A_CONV3S1: rows per TMA box when the NEXT tile of this CTA is prefetched into L2 (0 = off)
    // split weights, paired form (K == 2*ka, BN <= 128): a ring stage holds [A | W_hi | W_lo] of one k-block, the k loop runs
    // over ka only and every A tile is multiplied by both weight tiles - A crosses L2 -> shared memory once instead of twice
    int pair, pair_stages, k_lo;
    int epi;
    int round_out;            // fp32 output rounded to the nearest TF32 value
    int out_pair;             // 16-bit output rows [hi (N) | lo (N)], lo = residual of the rounding of hi
    int partial;              // split-K: raw fp32 partial sums to D + ks * part_stride, no bias / activation
    int64_t part_stride;
    const float* bias;
    int64_t b_gs;
    void* D;
    int64_t ldd, d_gs;
    const void* aux;
    int64_t aux_gs;
    int Hout_e, Wout_e;
    // weight-stationary mode: the whole [BN x K] weight slab of the current group stays in shared memory (loaded once per
    // group change), the ring carries A tiles only (ws_stages of them)
    int ws, ws_stages;
};
constexpr int MAX_STAGES = 12;     // barrier slots (weight-stationary rings are deeper than the A+B rings)

template <typename TO> __device__ __forceinline__ void store8(TO* p, const float* v);
template <> __device__ __forceinline__ void store8<float>(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
}
template <> __device__ __forceinline__ void store8<__half>(__half* p, const float* v) {
    // saturating: an out-of-range value becomes +-65504 instead of inf
    uint4 u;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.y) : "f"(v[3]), "f"(v[2]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.z) : "f"(v[5]), "f"(v[4]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.w) : "f"(v[7]), "f"(v[6]));
    *reinterpret_cast<uint4*>(p) = u;
}
// residuals of the 16-bit rounding of 8 values, themselves rounded to TO (the `lo` half of a pair output)
template <typename TO> __device__ __forceinline__ uint4 resid8(const float* v);
template <> __device__ __forceinline__ uint4 resid8<float>(const float*) { return make_uint4(0u, 0u, 0u, 0u); }
template <> __device__ __forceinline__ uint4 resid8<__nv_bfloat16>(const float* v) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
    __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]), b = __floats2bfloat162_rn(r[2], r[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(r[4], r[5]), d = __floats2bfloat162_rn(r[6], r[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    return u;
}
template <> __device__ __forceinline__ uint4 resid8<__half>(const float* v) {
    uint32_t h[4];
    uint4 u;
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
    float r[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        r[2 * i] = v[2 * i] - f.x; r[2 * i + 1] = v[2 * i + 1] - f.y;
    }
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.x) : "f"(r[1]), "f"(r[0]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.y) : "f"(r[3]), "f"(r[2]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.z) : "f"(r[5]), "f"(r[4]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.w) : "f"(r[7]), "f"(r[6]));
    return u;
}
// 8 consecutive elements of TO as loaded (the ADDUP epilogue keeps gathers in flight in this raw form)
template <typename TO> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void ld(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void unpack(float* v) const {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
};
template <> struct Raw8<__half> {
    uint4 u;
    __device__ __forceinline__ void ld(const __half* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void unpack(float* v) const {
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void ld(const float* p) { a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p) + 1); }
    __device__ __forceinline__ void unpack(float* v) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
};

struct TileCoord { int g, mt, nt, ks; };
__device__ __forceinline__ TileCoord decode_tile(int t, const TcParams& p) {
    TileCoord c;
    c.nt = t % p.n_tiles; t /= p.n_tiles;
    c.ks = t % p.ksplit;  t /= p.ksplit;
    c.mt = t % p.m_tiles;
    c.g = t / p.m_tiles;
    return c;
}

// ---- epilogue of one 128 x BN tile, one warp = 32 accumulator rows x BN/2 columns ------------------------------
// TMEM -> registers in 32-column chunks (the load of chunk c+1 is in flight while chunk c is processed), bias from a
// shared-memory copy of the tile's bias slice, activation, conversion, 16-byte st.shared into a swizzled staging tile
// [32 rows][SW bytes], and one TMA store per staged group: the async proxy writes whole 128-byte row segments and
// clips rows >= M, so the epilogue warps spend no instructions on global addressing.
template <int BN, typename TO, int EPI, bool ROUND, bool PAIR = false>
__device__ __forceinline__ void epi_tile(const TcParams& p, const CUtensorMap* tmD, const TileCoord& tc, uint32_t taddr,
                                         const float* sbias, uint8_t* stage, uint32_t tempty_bar, int lane, int q,
                                         int half) {
    constexpr int OUT_B = (int)sizeof(TO);
    constexpr int HALF = BN / 2;                                           // columns per warp
    constexpr int BOX_COLS = (HALF * OUT_B >= 128) ? 128 / OUT_B : HALF;   // columns per TMA-store box
    constexpr int SW = BOX_COLS * OUT_B;                                   // staged row bytes: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    constexpr int NCH = HALF / 32;                                         // 32-column chunks per warp
    constexpr int CPG = BOX_COLS / 32;                                     // chunks per store group
    static_assert(SW == 128 || SW == 64, "staging row");
    const int swz = (SW == 128) ? (lane & 7) : ((lane >> 1) & 3);
    const int m_w0 = tc.mt * BM + q * 32;                                  // first row of this warp
    const int n_base = tc.nt * BN + half * HALF;
    const int dz = p.partial ? tc.ks : tc.g;
    // ADDUP geometry of this thread's row
    int src00 = 0, src01 = 0, src10 = 0, src11 = 0;
    float wy0 = 0.f, wy1 = 0.f, wx0 = 0.f, wx1 = 0.f;
    const TO* aux = nullptr;
    bool row_ok = false;
    constexpr bool UPE = (EPI == EPI_RELU_ADDUP || EPI == EPI_ADDUP_RELU);      // bilinear-x2 source gathered per row
    if (EPI == EPI_ADD_RELU) {
        const int m = m_w0 + lane;
        row_ok = m < p.M;
        if (row_ok) aux = reinterpret_cast<const TO*>(p.aux) + (int64_t)tc.g * p.aux_gs + (int64_t)m * p.N;
    }
    if (UPE) {
        const int m = m_w0 + lane;
        row_ok = m < p.M;
        if (row_ok) {
            const int hw = p.Hout_e * p.Wout_e;
            const int im = m / hw, r = m - im * hw;
            const int y = r / p.Wout_e, x = r - y * p.Wout_e;
            const int hs = p.Hout_e >> 1, ws = p.Wout_e >> 1;
            const Up2Coef cy = up2_coef(y, hs), cx = up2_coef(x, ws);
            wy0 = cy.l0; wy1 = cy.l1; wx0 = cx.l0; wx1 = cx.l1;
            src00 = cy.i0 * ws + cx.i0; src01 = cy.i0 * ws + cx.i1;
            src10 = cy.i1 * ws + cx.i0; src11 = cy.i1 * ws + cx.i1;
            aux = reinterpret_cast<const TO*>(p.aux) + (int64_t)tc.g * p.aux_gs + (int64_t)im * hs * ws * p.N;
        }
    }
    uint32_t v[2][32];
    constexpr bool pair_out = PAIR && (OUT_B == 2);
    uint4 lo[pair_out ? CPG * 4 : 1];          // pair output: the residuals of one store group, stored after its `hi` box
    tc_ld32_issue(taddr, v[0]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t (&cur)[32] = v[c & 1];
        // ADDUP: the 4 bilinear sources of the chunk's columns are gathered (L2) while the TMEM load is in flight;
        // bf16 keeps the whole chunk in flight (16 x 16 B), fp32 one 8-column group at a time
        constexpr int PF = UPE ? ((OUT_B == 2) ? 4 : 1) : 1;
        Raw8<TO> ax[PF][4];
        if (UPE && PF == 4 && row_ok) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                const int col = n_base + c * 32 + gq * 8;
                ax[gq][0].ld(aux + (int64_t)src00 * p.N + col); ax[gq][1].ld(aux + (int64_t)src01 * p.N + col);
                ax[gq][2].ld(aux + (int64_t)src10 * p.N + col); ax[gq][3].ld(aux + (int64_t)src11 * p.N + col);
            }
        }
        tc_ld32_wait(cur);
        if (c + 1 < NCH) {
            tc_ld32_issue(taddr + (c + 1) * 32, v[(c + 1) & 1]);
        } else {
            // every accumulator column of this warp is in registers: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
        }
        const int sub = c % CPG;
        if (sub == 0) {
            // the staging tile is the source of the previous group's TMA store: wait until it has been read
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
        }
        const float* sb = sbias + half * HALF + c * 32;
        const int n0 = n_base + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            float x[8];
            const float4 b0 = *reinterpret_cast<const float4*>(sb + j);
            const float4 b1 = *reinterpret_cast<const float4*>(sb + j + 4);
            x[0] = __uint_as_float(cur[j + 0]) + b0.x; x[1] = __uint_as_float(cur[j + 1]) + b0.y;
            x[2] = __uint_as_float(cur[j + 2]) + b0.z; x[3] = __uint_as_float(cur[j + 3]) + b0.w;
            x[4] = __uint_as_float(cur[j + 4]) + b1.x; x[5] = __uint_as_float(cur[j + 5]) + b1.y;
            x[6] = __uint_as_float(cur[j + 6]) + b1.z; x[7] = __uint_as_float(cur[j + 7]) + b1.w;
            if (EPI == EPI_RELU || EPI == EPI_RELU_ADDUP) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
            } else if (EPI == EPI_GELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = gelu_erf(x[i]);
            }
            if (EPI == EPI_ADD_RELU) {
                if (row_ok) {
                    float a8[8];
                    Raw8<TO> r8;
                    r8.ld(aux + n0 + j);
                    r8.unpack(a8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] += a8[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
            }
            if (UPE) {
                if (row_ok) {
                    float a00[8], a01[8], a10[8], a11[8];
                    constexpr int PFI = (PF == 4) ? 1 : 0;
                    const int gq = PFI * (j / 8);
                    if (PF == 1) {
                        ax[0][0].ld(aux + (int64_t)src00 * p.N + n0 + j); ax[0][1].ld(aux + (int64_t)src01 * p.N + n0 + j);
                        ax[0][2].ld(aux + (int64_t)src10 * p.N + n0 + j); ax[0][3].ld(aux + (int64_t)src11 * p.N + n0 + j);
                    }
                    ax[gq][0].unpack(a00); ax[gq][1].unpack(a01); ax[gq][2].unpack(a10); ax[gq][3].unpack(a11);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float u = wy0 * (wx0 * a00[i] + wx1 * a01[i]) + wy1 * (wx0 * a10[i] + wx1 * a11[i]);
                        if (EPI == EPI_RELU_ADDUP) x[i] += fmaxf(u, 0.f);
                        else x[i] += u;
                    }
                }
                if (EPI == EPI_ADDUP_RELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
                }
            }
            if (ROUND) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = round_tf32(x[i]);
            }
            // staged row of this lane: 16-byte piece k lives at ((k ^ swz) * 16)
            if (OUT_B == 2) {
                const int k = sub * 4 + j / 8;
                store8<TO>(reinterpret_cast<TO*>(stage + lane * SW + ((k ^ swz) << 4)), x);
                if (pair_out) lo[pair_out ? k : 0] = resid8<TO>(x);
            } else {
                const int k = j / 4;
                *reinterpret_cast<float4*>(stage + lane * SW + ((k ^ swz) << 4)) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4*>(stage + lane * SW + (((k + 1) ^ swz) << 4)) = make_float4(x[4], x[5], x[6], x[7]);
            }
        }
        if (sub == CPG - 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_3d(tmD, smem_u32(stage), n0 - sub * 32, m_w0, dz);
            if (OUT_B == 2 && pair_out) {
                // the `lo` box of the same columns, N further along the row, through the same staging tile
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int k = 0; k < CPG * 4; ++k) *reinterpret_cast<uint4*>(stage + lane * SW + ((k ^ swz) << 4)) = lo[pair_out ? k : 0];
                fence_async_smem();
                __syncwarp();
                if (lane == 0) tma_store_3d(tmD, smem_u32(stage), p.N + n0 - sub * 32, m_w0, dz);
            }
        }
    }
}

template <int BN, typename TI, typename TO>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const TcParams p) {
    using C = TcCfg<BN>;
    constexpr int BK = ROW_BYTES / (int)sizeof(TI);      // elements per k-block
    constexpr bool TF32 = sizeof(TI) == 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 6);
    float* sbias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * MAX_STAGES;
    const uint32_t tfull0 = empty0 + 8 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t wfull = tempty0 + 16, wempty = wfull + 8;
    const int n_stages = p.ws ? p.ws_stages : p.pair ? p.pair_stages : C::STAGES;
    const uint32_t w_bytes = p.ws ? (uint32_t)p.kb_total * C::STAGE_B : 0u;      // resident weight slab in front of the ring
    const uint32_t stage_bytes = p.ws ? (uint32_t)C::STAGE_A : (uint32_t)(C::STAGE + (p.pair ? C::STAGE_B : 0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.groups * p.m_tiles * p.n_tiles * p.ksplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, EPI_WARPS); }
        mbar_init(wfull, 1); mbar_init(wempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmD)) : "memory");
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), (uint32_t)C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // after the TMEM allocation (see common.cuh)
    pdl_wait();         // everything above overlapped the previous kernel's tail; global memory from here on

    if (warp == 0) {
        // ================= TMA producer (the whole warp runs the loop, one elected lane issues: tc_ptx.cuh elect_one) =================
        {
            int stage = 0, phase = 0, w_group = -1, w_loads = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, p);
                const int kb0 = tc.ks * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                if (p.ws && tc.g != w_group) {
                    // new weight set: wait until every MMA that read the previous slab has completed, then reload it
                    if (w_loads > 0) mbar_wait(wempty, (w_loads - 1) & 1);
                    if (elect_one()) {
                        mbar_expect_tx(wfull, w_bytes);
                        for (int kb = 0; kb < p.kb_total; ++kb)
                            tma_load_3d(smem_base + kb * C::STAGE_B, &tmB, wfull, kb * BK, tc.nt * BN, tc.g);
                    }
                    __syncwarp();
                    w_group = tc.g;
                    ++w_loads;
                }
                // conv tile geometry (rows of all groups form one contiguous image list)
                int img0 = 0, oy0 = 0;
                if (p.amode != A_PLAIN) {
                    const long long gm0 = (long long)tc.g * p.M + (long long)tc.mt * BM;
                    img0 = (int)(gm0 / p.HWout);
                    oy0 = (int)(gm0 % p.HWout) / p.Wout;
                }
                if (p.pf_rows > 0 && tc.nt == 0 && tc.ks == 0) {
                    // experiment (off by default, see g_opt_conv_prefetch): warm L2 with the input rows of this CTA's NEXT tile
                    const int tn = t + gridDim.x;
                    if (tn < total_tiles) {
                        const TileCoord nc = decode_tile(tn, p);
                        const long long gm = (long long)nc.g * p.M + (long long)nc.mt * BM;
                        const int nimg = (int)(gm / p.HWout), noy = (int)(gm % p.HWout) / p.Wout;
                        if (elect_one()) {
                            for (int ci = 0; ci < p.Cin; ci += BK) {
                                tma_prefetch_4d(&tmA, ci, 0, noy - 1, nimg);
                                tma_prefetch_4d(&tmA, ci, 0, noy - 1 + p.pf_rows, nimg);
                            }
                        }
                        __syncwarp();
                    }
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t full = full0 + 8 * stage;
                    const uint32_t sa = smem_base + w_bytes + stage * stage_bytes, sb = sa + C::STAGE_A;
                    const int k = kb * BK;
                    const int k_a = (k >= p.ka) ? k - p.ka : k;      // split weights: the second half of K re-reads A
                    // A-tile coordinates of this k-block (c4 unused by the 4-D modes)
                    int c0, c1, c2, c3, c4 = 0;
                    if (p.amode == A_PLAIN) {
                        c0 = k_a % p.kblk; c1 = tc.mt * BM; c2 = k_a / p.kblk; c3 = tc.g;
                    } else if (p.amode == A_CONV3S1) {
                        // tap (ky,kx) of the tile's output pixels = the same box shifted by (ky-1, kx-1); the zero padding is
                        // the TMA unit's out-of-bounds fill (x = -1 / W, y = -1 / H of each image)
                        const int tap = k_a / p.Cin, ci = k_a - tap * p.Cin;
                        const int ky = tap / 3, kx = tap - ky * 3;
                        c0 = ci; c1 = kx - 1; c2 = oy0 + ky - 1; c3 = img0;
                    } else {
                        int tap = k_a / p.Cin;
                        const int ci = k_a - tap * p.Cin;
                        if (p.tap_fixed >= 0) tap = p.tap_fixed;
                        const int ky = tap / 3, kx = tap - ky * 3;
                        // input pixel (2*oy + ky - 1, 2*ox + kx - 1) = pair index (oy + dy, ox + dx), parity (hp, wp)
                        const int wp = (kx == 1) ? 0 : 1, dx = (kx == 0) ? -1 : 0;
                        const int hp = (ky == 1) ? 0 : 1, dy = (ky == 0) ? -1 : 0;
                        c0 = ci + wp * p.Cin; c1 = dx; c2 = hp; c3 = oy0 + dy; c4 = img0;
                    }
                    if (elect_one()) {
                        mbar_expect_tx(full, stage_bytes);
                        if (p.amode == A_CONV3S2) tma_load_5d(sa, &tmA, full, c0, c1, c2, c3, c4);
                        else tma_load_4d(sa, &tmA, full, c0, c1, c2, c3);
                        if (!p.ws) tma_load_3d(sb, &tmB, full, k, tc.nt * BN, tc.g);
                        if (p.pair) tma_load_3d(sb + C::STAGE_B, &tmB, full, k + p.k_lo, tc.nt * BN, tc.g);      // W_lo tile of the same k-block
                    }
                    __syncwarp();
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (the whole warp runs the loop, one elected lane issues) =================
        {
            constexpr uint32_t idesc = make_idesc_fmt(BM, BN, TF32 ? 2u : (std::is_same<TI, __half>::value ? 0u : 1u));
            int stage = 0, phase = 0, it = 0, w_group = -1, w_loads = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const TileCoord tc = decode_tile(t, p);
                const int kb0 = tc.ks * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                const int acc = it & 1, acc_phase = (it >> 1) & 1;
                if (p.ws && tc.g != w_group) {
                    mbar_wait(wfull, w_loads & 1);
                    w_group = tc.g;
                    ++w_loads;
                }
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + w_bytes + stage * stage_bytes;
                    const uint32_t sb = p.ws ? smem_base + kb * C::STAGE_B : sa + C::STAGE_A;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
                    const uint64_t dl = make_smem_desc(sb + C::STAGE_B);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            // one MMA consumes 32 B of K (16 bf16 / 8 tf32): +32 B inside the 128 B swizzle row (encoded >> 4)
                            tc_mma<TF32>(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                        }
                        if (p.pair) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) tc_mma<TF32>(tmem_d, da + 2 * kk, dl + 2 * kk, idesc, 1u);
                        }
                        tc_commit(empty0 + 8 * stage);
                    }
                    __syncwarp();
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
                // last tile of this CTA that uses the resident slab: release it once these MMAs have completed
                const int tn = t + gridDim.x;
                const bool release_w = p.ws && tn < total_tiles && decode_tile(tn, p).g != tc.g;
                if (elect_one()) {
                    tc_commit(tfull0 + 8 * acc);
                    if (release_w) tc_commit(wempty);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue =================
        const int e = warp - 2;
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        const int half = e >> 2;           // column half
        uint8_t* stage = smem + C::OFF_STAGING + e * 4096;                 // [32 rows][<= 128 B], private to this warp
        int it = 0, bias_key = -1;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const TileCoord tc = decode_tile(t, p);
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            // bias slice of this (group, n-tile), shared by the epilogue warps; zeros for split-K partials / no bias
            const int key = tc.g * p.n_tiles + tc.nt;
            if (key != bias_key) {
                bias_key = key;
                named_bar_sync_1<EPI_WARPS * 32>();            // nobody still reads the previous slice
                const int i = e * 32 + lane;
                if (i < BN) sbias[i] = (p.bias && !p.partial) ? __ldg(p.bias + (int64_t)tc.g * p.b_gs + tc.nt * BN + i) : 0.f;
                named_bar_sync_1<EPI_WARPS * 32>();
            }
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * (BN / 2);
            const uint32_t tempty = tempty0 + 8 * acc;
            const int epi = p.partial ? (int)EPI_NONE : p.epi;
            const bool rnd = sizeof(TO) == 4 && p.round_out && !p.partial;
#define EGR_EPI(E_, R_) epi_tile<BN, TO, E_, R_>(p, &tmD, tc, taddr, sbias, stage, tempty, lane, q, half)
            if (sizeof(TO) == 4 && rnd) {
                if (epi == EPI_RELU) EGR_EPI(EPI_RELU, (sizeof(TO) == 4));
                else if (epi == EPI_GELU) EGR_EPI(EPI_GELU, (sizeof(TO) == 4));
                else if (epi == EPI_RELU_ADDUP) EGR_EPI(EPI_RELU_ADDUP, (sizeof(TO) == 4));
                else EGR_EPI(EPI_NONE, (sizeof(TO) == 4));
            } else if (std::is_same<TO, __half>::value && p.out_pair) {
                // pair output [hi | lo] (split activations between two fp16 stages): the epilogues that feed such stages
                if (epi == EPI_RELU) epi_tile<BN, TO, EPI_RELU, false, true>(p, &tmD, tc, taddr, sbias, stage, tempty, lane, q, half);
                else if (epi == EPI_RELU_ADDUP) epi_tile<BN, TO, EPI_RELU_ADDUP, false, true>(p, &tmD, tc, taddr, sbias, stage, tempty, lane, q, half);
                else epi_tile<BN, TO, EPI_NONE, false, true>(p, &tmD, tc, taddr, sbias, stage, tempty, lane, q, half);
            } else {
                if (epi == EPI_RELU) EGR_EPI(EPI_RELU, false);
                else if (epi == EPI_GELU) EGR_EPI(EPI_GELU, false);
                else if (epi == EPI_RELU_ADDUP) EGR_EPI(EPI_RELU_ADDUP, false);
                else if (sizeof(TO) == 2 && epi == EPI_ADD_RELU) EGR_EPI(EPI_ADD_RELU, false);          // backbone stages: 16-bit outputs only
                else if (sizeof(TO) == 2 && epi == EPI_ADDUP_RELU) EGR_EPI(EPI_ADDUP_RELU, false);
                else EGR_EPI(EPI_NONE, false);
            }
#undef EGR_EPI
        }
        if (lane == 0) tma_store_wait_all();      // the staging tiles must outlive the last bulk stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)C::TMEM_COLS);
    }
}

// split-K finalize: out = epi(sum_s part[s] + bias)
template <typename TO>
__global__ void splitk_finalize_kernel(const float* part, int64_t part_stride, int ksplit, const float* bias, TO* out,
                                       int64_t total, int N, int epi, int round_out) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
    for (int k = 0; k < ksplit; ++k) s += part[(int64_t)k * part_stride + i];
    if (bias) s += bias[i % N];
    if (epi == EPI_RELU) s = fmaxf(s, 0.f);
    else if (epi == EPI_GELU) s = gelu_erf(s);
    if (sizeof(TO) == 4 && round_out) s = round_tf32(s);
    ActT<TO>::st(out + i, s);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_tc_mu;
bool g_tc_ready = false;
// split-K partial sums: one scratch buffer per launch stream (forwards of different streams may overlap)
std::unordered_map<cudaStream_t, float*> g_splitk_scratch;
constexpr int64_t SPLITK_SCRATCH_FLOATS = SPLITK_SCRATCH_BYTES / 4;
thread_local float* t_scratch = nullptr;

int encode(CUtensorMap* tm, int dt, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box, const char* what, bool swizzle64 = false, bool no_swizzle = false) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType tdt = dt == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                  : dt == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = g_encode(tm, tdt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(EGR_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)",
                    what, (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                    (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return EGR_OK;
}

template <int BN, typename TI, typename TO>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const TcParams& p, int grid, cudaStream_t st) {
    EGR_LAUNCH((gemm_tc_kernel<BN, TI, TO>), grid, TC_THREADS, TcCfg<BN>::SMEM, st, tmA, tmB, tmD, p);
    return EGR_OK;
}
// instantiated (operand, output) pairs: bf16 | fp16 -> bf16 | fp16 | fp32, fp32 (TF32) -> fp32
template <int BN, typename TI>
int launch_tc_out(int out_dt, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const TcParams& p, int grid, cudaStream_t st) {
    if (out_dt == DT_BF16) return launch_tc<BN, TI, __nv_bfloat16>(tmA, tmB, tmD, p, grid, st);
    if (out_dt == DT_F16) return launch_tc<BN, TI, __half>(tmA, tmB, tmD, p, grid, st);
    return launch_tc<BN, TI, float>(tmA, tmB, tmD, p, grid, st);
}
template <int BN>
int launch_tc_bn(int in_dt, int out_dt, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const TcParams& p, int grid, cudaStream_t st) {
    if (in_dt == DT_F32) return launch_tc<BN, float, float>(tmA, tmB, tmD, p, grid, st);
    if (in_dt == DT_BF16) return launch_tc_out<BN, __nv_bfloat16>(out_dt, tmA, tmB, tmD, p, grid, st);
    return launch_tc_out<BN, __half>(out_dt, tmA, tmB, tmD, p, grid, st);
}
int launch_tc_any(int bn, int in_dt, int out_dt, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const TcParams& p, int grid, cudaStream_t st) {
    if (bn == 256) return launch_tc_bn<256>(in_dt, out_dt, tmA, tmB, tmD, p, grid, st);
    if (bn == 128) return launch_tc_bn<128>(in_dt, out_dt, tmA, tmB, tmD, p, grid, st);
    return launch_tc_bn<64>(in_dt, out_dt, tmA, tmB, tmD, p, grid, st);
}
template <int BN, typename TI, typename TO>
int set_smem_attr() {
    EGR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM));
    return EGR_OK;
}
template <int BN>
int set_smem_attr_bn() {
    int rc;
    if ((rc = set_smem_attr<BN, __nv_bfloat16, float>()) || (rc = set_smem_attr<BN, __nv_bfloat16, __nv_bfloat16>()) ||
        (rc = set_smem_attr<BN, float, float>()) || (rc = set_smem_attr<BN, __half, float>()) || (rc = set_smem_attr<BN, __half, __half>()) ||
        (rc = set_smem_attr<BN, __half, __nv_bfloat16>()) || (rc = set_smem_attr<BN, __nv_bfloat16, __half>()))
        return rc;
    return EGR_OK;
}

}  // namespace

void gemm_tc_set_scratch(float* scratch) { t_scratch = scratch; }

// tiled tensor map (SWIZZLE_128B, or unswizzled rows) for the other TMA kernels of the library (head_tail_mma.cu)
int tc_encode_tiled(CUtensorMap* tm, int dt, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const char* what, bool no_swizzle) {
    if (int rc = gemm_tc_init()) return rc;
    cuuint64_t d[5], s[5];
    cuuint32_t b[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; if (i + 1 < rank) s[i] = strides_bytes[i]; }
    return encode(tm, dt, base, rank, d, s, b, what, false, no_swizzle);
}

int gemm_tc_init() {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    if (g_tc_ready) return EGR_OK;
    if (int rc = require_device()) return rc;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    EGR_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EGR_CHECK(fn && qres == cudaDriverEntryPointSuccess, EGR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    int rc;
    if ((rc = set_smem_attr_bn<64>()) || (rc = set_smem_attr_bn<128>()) || (rc = set_smem_attr_bn<256>())) return rc;
    g_tc_ready = true;
    return EGR_OK;
}

int gemm_tc(const GemmDesc& d, int in_dt, int out_dt_req, cudaStream_t st) {
    if (!g_tc_ready) {
        if (int rc = gemm_tc_init()) return rc;
    }
    EGR_CHECK(d.M > 0 && d.N > 0 && d.K > 0 && d.groups > 0, EGR_ERR_INVALID, "gemm_tc: empty problem %d %d %d", d.M, d.N, d.K);
    EGR_CHECK((in_dt == DT_F32 || in_dt == DT_BF16 || in_dt == DT_F16) &&
              (out_dt_req == DT_F32 || (in_dt != DT_F32 && (out_dt_req == DT_BF16 || out_dt_req == DT_F16))),
              EGR_ERR_UNSUPPORTED, "gemm_tc: operand / output types %d -> %d (TF32 operands give fp32; 16-bit operands give fp32, bf16 or fp16)", in_dt, out_dt_req);
    const bool f32 = in_dt == DT_F32;
    const int Ka = d.ka > 0 ? d.ka : d.K;       // A's own K (split operands: ka < K <= 2 * ka, the tail of K re-reads A from 0)
    EGR_CHECK(d.ka == 0 || (d.ka < d.K && d.K <= 2 * d.ka), EGR_ERR_INVALID, "gemm_tc: split operands need ka < K <= 2 * ka (K=%d ka=%d)", d.K, d.ka);
    const int ES = f32 ? 4 : 2;                 // operand element size
    const int BK = ROW_BYTES / ES;              // elements per k-block
    const int AL = 16 / ES;                     // elements per 16 bytes (TMA stride granularity)
    EGR_CHECK(d.K % BK == 0 && Ka % BK == 0, EGR_ERR_UNSUPPORTED, "gemm_tc: K=%d must be a multiple of %d", d.K, BK);
    EGR_CHECK(d.N % 64 == 0, EGR_ERR_UNSUPPORTED, "gemm_tc: N=%d must be a multiple of 64", d.N);
    EGR_CHECK((reinterpret_cast<uintptr_t>(d.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.W) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(d.D) & 15) == 0, EGR_ERR_INVALID, "gemm_tc: operands must be 16-byte aligned");
    EGR_CHECK(d.ldd % 8 == 0 && d.d_gs % 8 == 0, EGR_ERR_UNSUPPORTED, "gemm_tc: ldd / d_gs must be multiples of 8 elements");
    if (d.epi == EPI_RELU_ADDUP || d.epi == EPI_ADDUP_RELU) EGR_CHECK(d.aux && d.Hout > 0 && d.Wout > 0, EGR_ERR_INVALID, "gemm_tc: ADDUP needs aux + geometry");
    if (d.epi == EPI_ADD_RELU || d.epi == EPI_ADDUP_RELU)
        EGR_CHECK(d.aux && out_dt_req != DT_F32, EGR_ERR_UNSUPPORTED, "gemm_tc: the residual epilogues need aux and a 16-bit output");

    EGR_CHECK(!d.out_pair || (out_dt_req == DT_F16 && in_dt == DT_F16 && d.ldd >= 2 * (int64_t)d.N &&
                              (d.epi == EPI_NONE || d.epi == EPI_RELU || d.epi == EPI_RELU_ADDUP)), EGR_ERR_UNSUPPORTED,
              "gemm_tc: pair output needs fp16 operands and output, ldd >= 2N and epilogue none / ReLU / ReLU + up2");
    const int nsm = sm_count();
    const int m_tiles = ceil_div(d.M, BM);
    const int kb_total = d.K / BK;
    const bool can_split = d.epi != EPI_RELU_ADDUP && d.epi != EPI_ADD_RELU && d.epi != EPI_ADDUP_RELU && d.groups == 1 && d.ldd == d.N && kb_total >= 32;
    auto split_of = [&](int64_t base_tiles) {      // deterministic split-K factor for a skinny problem (1 = none)
        if (!can_split || base_tiles * 4 > nsm) return 1;
        int ks = (int)(nsm / base_tiles);
        if (ks > kb_total / 8) ks = kb_total / 8;
        while (ks > 1 && (int64_t)d.M * d.N * ks > SPLITK_SCRATCH_FLOATS) --ks;
        if (ks < 1) ks = 1;
        const int per = ceil_div(kb_total, ks);
        return ceil_div(kb_total, per);
    };
    // tile width: the widest BN whose tile list fills the SMs; for problems too small for that, the BN whose
    // (tiles x split-K) CTA count is largest, wider first (fewer re-reads of A, more weight bytes in flight per SM)
    int bn = 0;
    const int cands[3] = {256, 128, 64};
    for (int c : cands) {
        if (d.N % c) continue;
        if ((int64_t)d.groups * m_tiles * (d.N / c) >= nsm) { bn = c; break; }
    }
    if (!bn) {
        int64_t best = -1;
        for (int c : cands) {
            if (d.N % c) continue;
            const int64_t base = (int64_t)d.groups * m_tiles * (d.N / c);
            int64_t ctas = base * split_of(base);
            if (ctas > nsm) ctas = nsm;
            if (ctas > best) { best = ctas; bn = c; }
        }
    }
    TcParams p{};
    p.M = d.M; p.N = d.N; p.K = d.K; p.ka = Ka;
    p.m_tiles = m_tiles; p.n_tiles = d.N / bn; p.groups = d.groups;
    p.kb_total = kb_total;
    p.ksplit = 1; p.kb_per_split = p.kb_total;
    p.amode = d.amode; p.epi = d.epi; p.round_out = d.round_tf32; p.out_pair = d.out_pair;
    p.bias = d.bias; p.b_gs = d.b_gs;
    p.D = d.D; p.ldd = d.ldd; p.d_gs = d.d_gs;
    p.aux = d.aux; p.aux_gs = d.aux_gs; p.Hout_e = d.Hout; p.Wout_e = d.Wout;

    // deterministic split-K for skinny problems (pose3d Linear(32768 -> 2048) at small batch): partial sums go to a
    // library-owned scratch [ks][M][N], a finalize kernel adds them in a fixed order and applies bias + activation
    const int64_t base_tiles = (int64_t)d.groups * m_tiles * p.n_tiles;
    const int ks = split_of(base_tiles);
    if (ks > 1) {
        p.kb_per_split = ceil_div(p.kb_total, ks);
        p.ksplit = ks;
        p.partial = 1;
        p.part_stride = (int64_t)d.M * d.N;
        if (t_scratch) {
            p.D = t_scratch;
        } else {
            std::lock_guard<std::mutex> lk(g_tc_mu);
            float*& sc = g_splitk_scratch[st];
            if (!sc) EGR_CUDA_OK(cudaMalloc(&sc, SPLITK_SCRATCH_FLOATS * sizeof(float)));
            p.D = sc;
        }
    }

    // ---- tensor maps ----
    CUtensorMap tmA, tmB;
    int rc;
    if (d.amode == A_PLAIN) {
        const int kblk = d.kblk > 0 ? d.kblk : Ka;
        EGR_CHECK(kblk % BK == 0 && Ka % kblk == 0, EGR_ERR_UNSUPPORTED, "gemm_tc: kblk=%d", kblk);
        EGR_CHECK(d.lda % AL == 0 && d.a_gs % AL == 0 && d.kblk_stride % AL == 0, EGR_ERR_UNSUPPORTED, "gemm_tc: A strides must be multiples of 16 bytes");
        p.kblk = kblk;
        const cuuint64_t dims[4] = {(cuuint64_t)kblk, (cuuint64_t)d.M, (cuuint64_t)(Ka / kblk), (cuuint64_t)d.groups};
        // strides of size-1 dims are irrelevant but must be valid (multiple of 16 B)
        const cuuint64_t str[3] = {(cuuint64_t)d.lda * ES, (cuuint64_t)(Ka / kblk > 1 ? d.kblk_stride : d.lda) * ES,
                                   (cuuint64_t)(d.groups > 1 ? d.a_gs : d.lda) * ES};
        const cuuint32_t box[4] = {(cuuint32_t)BK, BM, 1, 1};
        if ((rc = encode(&tmA, in_dt, d.A, 4, dims, str, box, "A"))) return rc;
    } else if (d.amode == A_CONV3S1) {
        const int H = d.Hin, W = d.Win, HW = H * W;
        EGR_CHECK(d.Cin % BK == 0 && Ka == 9 * d.Cin, EGR_ERR_UNSUPPORTED, "gemm_tc: conv3s1 geometry Cin=%d K=%d", d.Cin, d.K);
        EGR_CHECK(W <= 256 && BM % W == 0 && (HW % BM == 0 || BM % HW == 0), EGR_ERR_UNSUPPORTED, "gemm_tc: conv3s1 output %dx%d does not tile by %d rows", H, W, BM);
        EGR_CHECK(d.M % HW == 0, EGR_ERR_INVALID, "gemm_tc: conv M=%d is not a whole number of %dx%d images", d.M, H, W);
        EGR_CHECK(d.groups == 1 || (d.a_gs == (int64_t)d.M * d.Cin && d.M % BM == 0), EGR_ERR_UNSUPPORTED,
                  "gemm_tc: conv groups must be contiguous image blocks");
        const int bh = (HW >= BM) ? BM / W : H;
        const int bimg = (HW >= BM) ? 1 : BM / HW;
        p.Cin = d.Cin; p.Wout = W; p.HWout = HW; p.tap_fixed = -1;
        p.pf_rows = (g_opt_conv_prefetch && HW >= BM && bh >= 2) ? bh : 0;
        const int64_t n_img = (int64_t)d.groups * (d.M / HW);
        const cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
        const cuuint64_t str[3] = {(cuuint64_t)d.Cin * ES, (cuuint64_t)W * d.Cin * ES, (cuuint64_t)H * W * d.Cin * ES};
        const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)W, (cuuint32_t)bh, (cuuint32_t)bimg};
        if ((rc = encode(&tmA, in_dt, d.A, 4, dims, str, box, "A(conv3s1)"))) return rc;
    } else {
        const int Hout = d.Hin / 2, Wout = d.Win / 2, HW = Hout * Wout;
        EGR_CHECK(d.Cin % BK == 0 && (Ka == 9 * d.Cin || Ka == d.Cin) && d.Hin % 2 == 0 && d.Win % 2 == 0, EGR_ERR_UNSUPPORTED,
                  "gemm_tc: conv geometry Cin=%d K=%d", d.Cin, d.K);
        p.tap_fixed = (Ka == d.Cin) ? 4 : -1;        // K == Cin: centre tap only = 1x1 stride-2 conv
        EGR_CHECK(BM % Wout == 0 && (HW % BM == 0 || BM % HW == 0), EGR_ERR_UNSUPPORTED, "gemm_tc: conv output %dx%d does not tile by %d rows", Hout, Wout, BM);
        EGR_CHECK(d.M % HW == 0, EGR_ERR_INVALID, "gemm_tc: conv M=%d is not a whole number of %dx%d images", d.M, Hout, Wout);
        EGR_CHECK(d.groups == 1 || (d.a_gs == (int64_t)(d.M / HW) * d.Hin * d.Win * d.Cin && d.M % BM == 0), EGR_ERR_UNSUPPORTED,
                  "gemm_tc: conv groups must be contiguous image blocks");
        const int bh = (HW >= BM) ? BM / Wout : Hout;
        const int bimg = (HW >= BM) ? 1 : BM / HW;
        p.Cin = d.Cin; p.Wout = Wout; p.HWout = HW;
        const int64_t n_img = (int64_t)d.groups * (d.M / HW);
        const cuuint64_t dims[5] = {(cuuint64_t)2 * d.Cin, (cuuint64_t)Wout, 2, (cuuint64_t)Hout, (cuuint64_t)n_img};
        const cuuint64_t str[4] = {(cuuint64_t)2 * d.Cin * ES, (cuuint64_t)d.Win * d.Cin * ES, (cuuint64_t)2 * d.Win * d.Cin * ES,
                                   (cuuint64_t)d.Hin * d.Win * d.Cin * ES};
        const cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)Wout, 1, (cuuint32_t)bh, (cuuint32_t)bimg};
        if ((rc = encode(&tmA, in_dt, d.A, 5, dims, str, box, "A(conv3s2)"))) return rc;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)d.K, (cuuint64_t)d.N, (cuuint64_t)d.groups};
        const cuuint64_t str[2] = {(cuuint64_t)d.K * ES, (cuuint64_t)(d.groups > 1 ? d.w_gs : (int64_t)d.N * d.K) * ES};
        const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)bn, 1};
        if ((rc = encode(&tmB, in_dt, d.W, 3, dims, str, box, "W"))) return rc;
    }

    const int out_dt = p.partial ? (int)DT_F32 : out_dt_req;
    const bool out_bf16 = out_dt != DT_F32;          // 16-bit output
    CUtensorMap tmD;
    {
        // output [groups | ksplit][M][ldd]: one TMA-store box = 32 rows x (128 B, or the warp's BN/2 columns when narrower)
        const int OB = out_bf16 ? 2 : 4;
        const int box_cols = ((bn / 2) * OB >= 128) ? 128 / OB : bn / 2;
        const int64_t ldd = p.partial ? d.N : d.ldd;
        const int nz = p.partial ? p.ksplit : d.groups;
        const int64_t zs = p.partial ? p.part_stride : d.d_gs;
        EGR_CHECK((ldd * OB) % 16 == 0 && (nz == 1 || (zs * OB) % 16 == 0), EGR_ERR_UNSUPPORTED, "gemm_tc: D strides must be multiples of 16 bytes");
        const cuuint64_t dims[3] = {(cuuint64_t)d.N * (d.out_pair ? 2 : 1), (cuuint64_t)d.M, (cuuint64_t)nz};
        const cuuint64_t str[2] = {(cuuint64_t)ldd * OB, (cuuint64_t)(nz > 1 ? zs : (int64_t)d.M * ldd) * OB};
        const cuuint32_t box[3] = {(cuuint32_t)box_cols, 32, 1};
        if ((rc = encode(&tmD, out_dt, p.D, 3, dims, str, box, "D", box_cols * OB == 64))) return rc;
    }
    // split weights: pair W_hi / W_lo in one ring stage when the tile is narrow enough to keep >= 4 stages
    if (g_opt_pair && d.ka > 0 && d.K == 2 * d.ka && bn <= 128 && p.ksplit == 1) {
        p.pair = 1;
        p.pair_stages = SMEM_RING / (BM * ROW_BYTES + 2 * bn * ROW_BYTES);
        p.k_lo = Ka;
        p.kb_total = Ka / BK;
        p.kb_per_split = p.kb_total;
    }
    const int64_t total = (int64_t)d.groups * m_tiles * p.n_tiles * p.ksplit;
    const int grid = (int)(total < nsm ? total : nsm);
    // weight-stationary when the group's whole weight slab fits beside >= 4 A stages and every CTA walks several tiles:
    // the slab is fetched once per CTA (and group) instead of once per tile, and the ring becomes A-only and deeper
    {
        const int64_t w_bytes = (int64_t)p.kb_total * bn * ROW_BYTES;
        const int64_t a_stages = (SMEM_RING - w_bytes) / (BM * ROW_BYTES);
        const bool ws_ok = d.ka == 0 && !p.pair && (d.amode != A_PLAIN || p.kblk == d.K) && p.n_tiles == 1 && p.ksplit == 1 && a_stages >= 4 &&
                           total >= 2 * (int64_t)grid;
        if (ws_ok && (d.ws || (g_opt_ws && d.amode == A_PLAIN))) {
            p.ws = 1;
            p.ws_stages = (int)(a_stages < MAX_STAGES ? a_stages : MAX_STAGES);
        }
    }
    rc = launch_tc_any(bn, in_dt, out_dt, tmA, tmB, tmD, p, grid, st);
    if (rc) return rc;
    if (p.partial) {
        const int64_t tot = (int64_t)d.M * d.N;
        const int blocks = (int)ceil_div64(tot, 256);
        if (out_dt_req == DT_F16)
            EGR_LAUNCH(splitk_finalize_kernel<__half>, blocks, 256, 0, st, reinterpret_cast<const float*>(p.D), p.part_stride, p.ksplit, d.bias,
                       reinterpret_cast<__half*>(d.D), tot, d.N, d.epi, 0);
        else if (out_dt_req == DT_BF16)
            EGR_LAUNCH(splitk_finalize_kernel<__nv_bfloat16>, blocks, 256, 0, st, reinterpret_cast<const float*>(p.D), p.part_stride, p.ksplit, d.bias,
                       reinterpret_cast<__nv_bfloat16*>(d.D), tot, d.N, d.epi, 0);
        else
            EGR_LAUNCH(splitk_finalize_kernel<float>, blocks, 256, 0, st, reinterpret_cast<const float*>(p.D), p.part_stride, p.ksplit, d.bias,
                       reinterpret_cast<float*>(d.D), tot, d.N, d.epi, d.round_tf32);
    }
    return EGR_OK;
}

}  // namespace egr
