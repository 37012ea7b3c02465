// HBM-bound layout / resampling stages: NCHW->NHWC staging, bilinear x2 (align_corners=True) tails of the
// heatmap heads and of the refined-feature projection, 2x2 max-pool, and one-time weight repacking.
#include "layout_ops.cuh"

namespace egr {

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

int nchw_to_nhwc(const float* in, void* out, int B, int V, int C, int HW, int out_mode, cudaStream_t st) {
    EGR_CHECK(HW % 32 == 0 && C * 33 * 4 <= 48 * 1024, EGR_ERR_UNSUPPORTED, "nchw_to_nhwc: HW=%d C=%d", HW, C);
    dim3 grid(HW / 32, B * V);
    const size_t smem = sizeof(float) * C * 33;
    if (out_mode == 1) nchw_to_nhwc_kernel<__nv_bfloat16, false><<<grid, 256, smem, st>>>(in, (__nv_bfloat16*)out, B, V, C, HW);
    else if (out_mode == 2) nchw_to_nhwc_kernel<float, true><<<grid, 256, smem, st>>>(in, (float*)out, B, V, C, HW);
    else nchw_to_nhwc_kernel<float, false><<<grid, 256, smem, st>>>(in, (float*)out, B, V, C, HW);
    EGR_LAUNCHED();
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

int head_up_conv(const void* z, int z_bf16, const float* w, const float* bias, const int* wsel_host, int B, int G,
                 int Hs, int Ws, int C, int J, float* hm, int64_t hm_bs, int64_t hm_gs, void* hm_t,
                 cudaStream_t st) {
    EGR_CHECK(J <= HJ && (2 * Hs) % STRIP == 0 && G <= 4, EGR_ERR_UNSUPPORTED, "head_up_conv: J=%d Hs=%d G=%d", J, Hs, G);
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

int up2_relu_dual(const void* z, int z_bf16, int B, int G, int Hs, int Ws, int C, float* out_nchw, int64_t o_bs,
                  int64_t o_gs, void* out_nhwc, cudaStream_t st) {
    EGR_CHECK((2 * Hs) % STRIP == 0, EGR_ERR_UNSUPPORTED, "up2_relu_dual: geometry");
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
template <typename T>
__global__ void maxpool2_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t total, int H, int W, int C) {
    const int Ho = H >> 1, Wo = W >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t r = i / C;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int64_t img = r / Ho;
        const T* p = in + ((img * H + 2 * oy) * W + 2 * ox) * C + c;
        const float a = ActT<T>::ld(p), b = ActT<T>::ld(p + C), cc = ActT<T>::ld(p + (int64_t)W * C),
                    d = ActT<T>::ld(p + (int64_t)W * C + C);
        ActT<T>::st(out + i, fmaxf(fmaxf(a, b), fmaxf(cc, d)));
    }
}

int maxpool2_nhwc(const void* in, void* out, int is_bf16, int64_t n_img, int H, int W, int C, cudaStream_t st) {
    const int64_t total = n_img * (H / 2) * (W / 2) * C;
    if (total == 0) return EGR_OK;
    const int grid = (int)(ceil_div64(total, 256) < 148 * 32 ? ceil_div64(total, 256) : 148 * 32);
    if (is_bf16) maxpool2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, total, H, W, C);
    else maxpool2_kernel<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, total, H, W, C);
    EGR_LAUNCHED();
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
int cast_act(const float* in, void* out, int out_bf16, int64_t n, cudaStream_t st) {
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
