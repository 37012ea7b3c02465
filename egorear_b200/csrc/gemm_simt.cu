// fp32 SIMT GEMM / implicit-GEMM conv: the reference-grade (EGR_PREC_FP32) dense stage.
// 128x64x16 block tile, 256 threads, 8x4 register micro-tile, fp32 FMA accumulation in k order.
#include "gemm.cuh"

namespace egr {

constexpr int SBM = 128, SBN = 64, SBK = 16, STHREADS = 256;

template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

template <typename TA, typename TO, int AMODE>
__global__ void __launch_bounds__(STHREADS)
gemm_simt_kernel(GemmDesc d) {
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Bs[SBK][SBN + 4];
    pdl_trigger();
    pdl_wait();

    const int g = blockIdx.z;
    const TA* A = reinterpret_cast<const TA*>(d.A) + (int64_t)g * d.a_gs;
    const float* W = reinterpret_cast<const float*>(d.W) + (int64_t)g * d.w_gs;
    const float* bias = d.bias ? d.bias + (int64_t)g * d.b_gs : nullptr;
    TO* D = reinterpret_cast<TO*>(d.D) + (int64_t)g * d.d_gs;

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;

    // A loader: row am, 8 consecutive k starting at ak
    const int am = tid >> 1, ak = (tid & 1) * 8;
    const int gm = m0 + am;
    const bool am_ok = gm < d.M;
    // conv geometry of this row
    int img = 0, oy = 0, ox = 0;
    if (AMODE == A_CONV3S2 && am_ok) {
        const int Wo = d.Win >> 1, Ho = d.Hin >> 1;
        img = gm / (Ho * Wo);
        const int r = gm - img * Ho * Wo;
        oy = r / Wo;
        ox = r - oy * Wo;
    }
    // B loader: row bn, 4 consecutive k starting at bk
    const int bn = tid >> 2, bk = (tid & 3) * 4;
    const bool bn_ok = (n0 + bn) < d.N;
    const float* Wrow = W + (int64_t)(n0 + bn) * d.K;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < d.K; k0 += SBK) {
        // ---- load A (8 values) ----
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (am_ok) {
            const int k = k0 + ak;
            if (AMODE == A_PLAIN) {
                const TA* p;
                if (d.kblk > 0) p = A + (int64_t)(k / d.kblk) * d.kblk_stride + (int64_t)gm * d.lda + (k % d.kblk);
                else p = A + (int64_t)gm * d.lda + k;
                a0 = ld4<TA>(p);
                a1 = ld4<TA>(p + 4);
            } else {
                const int tap = k / d.Cin, ci = k - tap * d.Cin;
                const int ky = tap / 3, kx = tap - ky * 3;
                const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
                if (iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win) {
                    const TA* p = A + (((int64_t)img * d.Hin + iy) * d.Win + ix) * d.Cin + ci;
                    a0 = ld4<TA>(p);
                    a1 = ld4<TA>(p + 4);
                }
            }
        }
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bn_ok) b0 = *reinterpret_cast<const float4*>(Wrow + k0 + bk);
        __syncthreads();
        As[ak + 0][am] = a0.x; As[ak + 1][am] = a0.y; As[ak + 2][am] = a0.z; As[ak + 3][am] = a0.w;
        As[ak + 4][am] = a1.x; As[ak + 5][am] = a1.y; As[ak + 6][am] = a1.z; As[ak + 7][am] = a1.w;
        Bs[bk + 0][bn] = b0.x; Bs[bk + 1][bn] = b0.y; Bs[bk + 2][bn] = b0.z; Bs[bk + 3][bn] = b0.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            const float4 av0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 av1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[8] = {av0.x, av0.y, av0.z, av0.w, av1.x, av1.y, av1.z, av1.w};
            const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

    // ---- epilogue ----
    const int n = n0 + tx * 4;
    if (n >= d.N) return;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n + j < d.N) bv[j] = bias[n + j];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= d.M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x = acc[i][j] + bv[j];
            if (d.epi == EPI_RELU || d.epi == EPI_RELU_ADDUP) x = fmaxf(x, 0.f);
            else if (d.epi == EPI_GELU) x = gelu_erf(x);
            v[j] = x;
        }
        if (d.epi == EPI_RELU_ADDUP) {
            // + relu(up2(aux)) at this output position; aux is [img][(Hout/2)*(Wout/2)][N], same dtype as D
            const TO* aux = reinterpret_cast<const TO*>(d.aux) + (int64_t)g * d.aux_gs;
            const int hw = d.Hout * d.Wout;
            const int im = m / hw, r = m - im * hw;
            const int y = r / d.Wout, x = r - y * d.Wout;
            const int hs = d.Hout >> 1, ws = d.Wout >> 1;
            const Up2Coef cy = up2_coef(y, hs), cx = up2_coef(x, ws);
            const TO* base = aux + (int64_t)im * hs * ws * d.N + n;
            const float4 v00 = ld4<TO>(base + (int64_t)(cy.i0 * ws + cx.i0) * d.N);
            const float4 v01 = ld4<TO>(base + (int64_t)(cy.i0 * ws + cx.i1) * d.N);
            const float4 v10 = ld4<TO>(base + (int64_t)(cy.i1 * ws + cx.i0) * d.N);
            const float4 v11 = ld4<TO>(base + (int64_t)(cy.i1 * ws + cx.i1) * d.N);
            const float u0 = cy.l0 * (cx.l0 * v00.x + cx.l1 * v01.x) + cy.l1 * (cx.l0 * v10.x + cx.l1 * v11.x);
            const float u1 = cy.l0 * (cx.l0 * v00.y + cx.l1 * v01.y) + cy.l1 * (cx.l0 * v10.y + cx.l1 * v11.y);
            const float u2 = cy.l0 * (cx.l0 * v00.z + cx.l1 * v01.z) + cy.l1 * (cx.l0 * v10.z + cx.l1 * v11.z);
            const float u3 = cy.l0 * (cx.l0 * v00.w + cx.l1 * v01.w) + cy.l1 * (cx.l0 * v10.w + cx.l1 * v11.w);
            v[0] += fmaxf(u0, 0.f); v[1] += fmaxf(u1, 0.f); v[2] += fmaxf(u2, 0.f); v[3] += fmaxf(u3, 0.f);
        }
        TO* dst = D + (int64_t)m * d.ldd + n;
        if (n + 3 < d.N) {
            st4<TO>(dst, make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < d.N) ActT<TO>::st(dst + j, v[j]);
        }
    }
}

template <typename TA, typename TO>
static int launch_simt(const GemmDesc& d, cudaStream_t st) {
    dim3 grid(ceil_div(d.M, SBM), ceil_div(d.N, SBN), d.groups);
    if (d.amode == A_PLAIN) EGR_LAUNCH((gemm_simt_kernel<TA, TO, A_PLAIN>), grid, STHREADS, 0, st, d);
    else EGR_LAUNCH((gemm_simt_kernel<TA, TO, A_CONV3S2>), grid, STHREADS, 0, st, d);
    return EGR_OK;
}

int gemm_simt(const GemmDesc& d, int a_is_bf16, int d_is_bf16, cudaStream_t st) {
    EGR_CHECK(d.M > 0 && d.N > 0 && d.K > 0, EGR_ERR_INVALID, "gemm: empty problem %d %d %d", d.M, d.N, d.K);
    EGR_CHECK(d.K % SBK == 0, EGR_ERR_UNSUPPORTED, "gemm_simt: K=%d must be a multiple of %d", d.K, SBK);
    EGR_CHECK((d.amode == A_PLAIN || d.amode == A_CONV3S2) && d.epi <= EPI_RELU_ADDUP && d.ka == 0 &&
              !(d.amode == A_CONV3S2 && d.K == d.Cin), EGR_ERR_UNSUPPORTED,
              "gemm_simt: amode %d / epi %d / split operands exist on the tensor-core path only", d.amode, d.epi);
    EGR_CHECK(d.N % 4 == 0 || d.ldd % 4 != 0 || true, EGR_ERR_UNSUPPORTED, "gemm_simt: N");
    if (d.amode == A_CONV3S2)
        EGR_CHECK(d.Cin % SBK == 0 && d.K == 9 * d.Cin && d.Hin % 2 == 0 && d.Win % 2 == 0, EGR_ERR_UNSUPPORTED,
                  "gemm_simt: conv geometry Cin=%d K=%d", d.Cin, d.K);
    if (d.amode == A_PLAIN && d.kblk > 0)
        EGR_CHECK(d.kblk % SBK == 0 && d.K % d.kblk == 0, EGR_ERR_UNSUPPORTED, "gemm_simt: kblk=%d", d.kblk);
    if (d.epi == EPI_RELU_ADDUP) EGR_CHECK(d.aux && d.N % 4 == 0, EGR_ERR_INVALID, "gemm_simt: ADDUP needs aux");
    if (a_is_bf16 && d_is_bf16) return launch_simt<__nv_bfloat16, __nv_bfloat16>(d, st);
    if (a_is_bf16 && !d_is_bf16) return launch_simt<__nv_bfloat16, float>(d, st);
    if (!a_is_bf16 && d_is_bf16) return launch_simt<float, __nv_bfloat16>(d, st);
    return launch_simt<float, float>(d, st);
}

}  // namespace egr
