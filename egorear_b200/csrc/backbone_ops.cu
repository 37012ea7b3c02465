// Backbone-side kernels around the tcgen05 conv stages (ResNet18 + EfficientFPN, models/backbones/resnet.py):
//   stem_im2col   NCHW fp32 image -> [pixel][7*7*3 padded to 192] 16-bit rows (A operand of the 7x7 stride-2 stem GEMM)
//   maxpool3s2    nn.MaxPool2d(3, 2, 1) on channels-last 16-bit maps
//   nhwc_to_nchw  stride-32 map [V][B][64][512] 16-bit -> [B][V][512][8][8] fp32 (the `backbone_feat_bottom` the jqa query reads)
//   bn_fold       eval-mode BatchNorm folded into the preceding conv (prepack): w' = w * g / sqrt(var + eps), b' = beta - mean * g / sqrt(..)
#include "backbone_ops.cuh"

namespace egr {

namespace {

template <typename T16> __device__ __forceinline__ uint32_t bb_pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t bb_pack2<__nv_bfloat16>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <> __device__ __forceinline__ uint32_t bb_pack2<__half>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// one thread = 8 consecutive k of one output pixel (16 bytes); k = (ky*7 + kx)*3 + c, zero for k >= 147 and for taps in the padding
template <typename T16>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, T16* __restrict__ out, int B, int V, int H, int W) {
    pdl_trigger();
    pdl_wait();
    const int Ho = H >> 1, Wo = W >> 1;
    const int64_t total = (int64_t)V * B * Ho * Wo * (STEM_K / 8);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k8 = (int)(i % (STEM_K / 8));
        int64_t r = i / (STEM_K / 8);
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho); r /= Ho;
        const int b = (int)(r % B);
        const int v = (int)(r / B);
        const float* src = img + ((int64_t)b * V + v) * 3 * H * W;
        float val[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k8 * 8 + e;
            float x = 0.f;
            if (k < 147) {
                const int tap = k / 3, c = k - tap * 3;
                const int ky = tap / 7, kx = tap - ky * 7;
                const int iy = 2 * oy + ky - 3, ix = 2 * ox + kx - 3;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) x = __ldg(src + ((int64_t)c * H + iy) * W + ix);
            }
            val[e] = x;
        }
        uint4 u;
        u.x = bb_pack2<T16>(val[0], val[1]); u.y = bb_pack2<T16>(val[2], val[3]);
        u.z = bb_pack2<T16>(val[4], val[5]); u.w = bb_pack2<T16>(val[6], val[7]);
        reinterpret_cast<uint4*>(out)[i] = u;
    }
}

template <typename T16> __device__ __forceinline__ uint32_t bb_max2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t bb_max2<__nv_bfloat16>(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t bb_max2<__half>(uint32_t a, uint32_t b) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// MaxPool2d(kernel 3, stride 2, padding 1): padding is -inf, i.e. out-of-map taps are skipped; thread = (pixel, 8 channels)
template <typename T16>
__global__ void __launch_bounds__(256)
maxpool3s2_kernel(const T16* __restrict__ in, T16* __restrict__ out, int64_t n_img, int H, int W, int C) {
    pdl_trigger();
    pdl_wait();
    const int Ho = H >> 1, Wo = W >> 1, CV = C >> 3;
    const int64_t total = n_img * Ho * Wo * CV;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        int64_t r = i / CV;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int64_t im = r / Ho;
        const uint4* base = reinterpret_cast<const uint4*>(in + im * (int64_t)H * W * C) + cv;
        uint4 m;
        bool first = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = 2 * oy + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int ix = 2 * ox + dx;
                if (ix < 0 || ix >= W) continue;
                const uint4 x = __ldg(base + ((int64_t)iy * W + ix) * CV);
                if (first) { m = x; first = false; }
                else { m.x = bb_max2<T16>(m.x, x.x); m.y = bb_max2<T16>(m.y, x.y); m.z = bb_max2<T16>(m.z, x.z); m.w = bb_max2<T16>(m.w, x.w); }
            }
        }
        reinterpret_cast<uint4*>(out)[i] = m;
    }
}

template <typename T16> __device__ __forceinline__ float2 bb_unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 bb_unpack2<__nv_bfloat16>(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 bb_unpack2<__half>(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
template <typename T16> __device__ __forceinline__ uint32_t bb_pack2_relu(float lo, float hi) {
    return bb_pack2<T16>(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
}

// thread = (output pixel, 8 channels): 16-byte accesses, consecutive lanes = consecutive channel groups of a pixel row
template <typename T16>
__global__ void __launch_bounds__(256)
up2_add_relu_kernel(const T16* __restrict__ hi, const T16* __restrict__ low, T16* __restrict__ out, int64_t n_img, int R, int C) {
    pdl_trigger();
    pdl_wait();
    const int CV = C >> 3, Rs = R >> 1;
    const int64_t total = n_img * R * R * CV;
    const float scale = (float)(Rs - 1) / (float)(R - 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        int64_t r = i / CV;
        const int x = (int)(r % R); r /= R;
        const int y = (int)(r % R);
        const int64_t im = r / R;
        const float sy = scale * (float)y, sx = scale * (float)x;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < Rs - 1 ? 1 : 0), x1 = x0 + (x0 < Rs - 1 ? 1 : 0);
        const float ly1 = sy - (float)y0, lx1 = sx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
        const uint4* lb = reinterpret_cast<const uint4*>(low + im * (int64_t)Rs * Rs * C) + cv;
        const uint4 a = __ldg(lb + ((int64_t)y0 * Rs + x0) * CV), b = __ldg(lb + ((int64_t)y0 * Rs + x1) * CV);
        const uint4 c = __ldg(lb + ((int64_t)y1 * Rs + x0) * CV), d = __ldg(lb + ((int64_t)y1 * Rs + x1) * CV);
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi) + i);
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cvv[4] = {c.x, c.y, c.z, c.w},
                       dv[4] = {d.x, d.y, d.z, d.w}, hv[4] = {h.x, h.y, h.z, h.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 fa = bb_unpack2<T16>(av[e]), fb = bb_unpack2<T16>(bv[e]), fc = bb_unpack2<T16>(cvv[e]), fd = bb_unpack2<T16>(dv[e]);
            const float2 fh = bb_unpack2<T16>(hv[e]);
            const float u0 = ly0 * (lx0 * fa.x + lx1 * fb.x) + ly1 * (lx0 * fc.x + lx1 * fd.x);
            const float u1 = ly0 * (lx0 * fa.y + lx1 * fb.y) + ly1 * (lx0 * fc.y + lx1 * fd.y);
            o[e] = bb_pack2_relu<T16>(fh.x + u0, fh.y + u1);
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

template <typename T16>
__global__ void __launch_bounds__(256)
bottom_to_nchw_kernel(const T16* __restrict__ in, float* __restrict__ out, int B, int V, int HW, int C) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = (int64_t)B * V * C * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        int64_t r = i / HW;
        const int c = (int)(r % C); r /= C;
        const int v = (int)(r % V);
        const int b = (int)(r / V);
        out[i] = ActT<T16>::ld(in + (((int64_t)v * B + b) * HW + p) * C + c);
    }
}

__global__ void bn_fold_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps, float* __restrict__ w_out,
                               float* __restrict__ b_out, int N, int K) {
    const int64_t total = (int64_t)N * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / K);
        const float s = gamma[n] / sqrtf(var[n] + eps);
        w_out[i] = w[i] * s;
        if (i - (int64_t)n * K == 0) b_out[n] = beta[n] - mean[n] * s;
    }
}

// stem weight [64][3][7][7] -> [64][(ky*7 + kx)*3 + c], zero padded to STEM_K
__global__ void stem_repack_kernel(const float* __restrict__ w, float* __restrict__ out, int N) {
    const int total = N * STEM_K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / STEM_K, k = i - n * STEM_K;
        float x = 0.f;
        if (k < 147) {
            const int tap = k / 3, c = k - tap * 3;
            x = w[(n * 3 + c) * 49 + tap];
        }
        out[i] = x;
    }
}

inline int grid_for(int64_t total) {
    const int64_t g = ceil_div64(total, 256);
    const int64_t cap = (int64_t)sm_count() * 32;
    return (int)(g < cap ? g : cap);
}

}  // namespace

int stem_im2col(const float* img, void* out, int dt, int B, int V, int H, int W, cudaStream_t st) {
    EGR_CHECK(H % 2 == 0 && W % 2 == 0 && (dt == 1 || dt == 2), EGR_ERR_UNSUPPORTED, "stem_im2col: H=%d W=%d dt=%d", H, W, dt);
    const int64_t total = (int64_t)V * B * (H / 2) * (W / 2) * (STEM_K / 8);
    if (dt == 2) EGR_LAUNCH(stem_im2col_kernel<__half>, grid_for(total), 256, 0, st, img, (__half*)out, B, V, H, W);
    else EGR_LAUNCH(stem_im2col_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, img, (__nv_bfloat16*)out, B, V, H, W);
    return EGR_OK;
}

int maxpool3s2_nhwc(const void* in, void* out, int dt, int64_t n_img, int H, int W, int C, cudaStream_t st) {
    EGR_CHECK(C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && (dt == 1 || dt == 2), EGR_ERR_UNSUPPORTED, "maxpool3s2: C=%d H=%d W=%d dt=%d", C, H, W, dt);
    const int64_t total = n_img * (H / 2) * (W / 2) * (C / 8);
    if (dt == 2) EGR_LAUNCH(maxpool3s2_kernel<__half>, grid_for(total), 256, 0, st, (const __half*)in, (__half*)out, n_img, H, W, C);
    else EGR_LAUNCH(maxpool3s2_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n_img, H, W, C);
    return EGR_OK;
}

int up2_add_relu_nhwc(const void* hi, const void* low, void* out, int dt, int64_t n_img, int R, int C, cudaStream_t st) {
    EGR_CHECK(C % 8 == 0 && R % 2 == 0 && (dt == 1 || dt == 2), EGR_ERR_UNSUPPORTED, "up2_add_relu: C=%d R=%d dt=%d", C, R, dt);
    const int64_t total = n_img * R * R * (C / 8);
    if (dt == 2) EGR_LAUNCH(up2_add_relu_kernel<__half>, grid_for(total), 256, 0, st, (const __half*)hi, (const __half*)low, (__half*)out, n_img, R, C);
    else EGR_LAUNCH(up2_add_relu_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, (const __nv_bfloat16*)hi, (const __nv_bfloat16*)low,
                    (__nv_bfloat16*)out, n_img, R, C);
    return EGR_OK;
}

int bottom_to_nchw(const void* in, float* out, int dt, int B, int V, int HW, int C, cudaStream_t st) {
    EGR_CHECK(dt == 1 || dt == 2, EGR_ERR_UNSUPPORTED, "bottom_to_nchw: dt=%d", dt);
    const int64_t total = (int64_t)B * V * C * HW;
    if (dt == 2) EGR_LAUNCH(bottom_to_nchw_kernel<__half>, grid_for(total), 256, 0, st, (const __half*)in, out, B, V, HW, C);
    else EGR_LAUNCH(bottom_to_nchw_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, (const __nv_bfloat16*)in, out, B, V, HW, C);
    return EGR_OK;
}

int bn_fold(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* w_out,
            float* b_out, int N, int K, cudaStream_t st) {
    bn_fold_kernel<<<grid_for((int64_t)N * K), 256, 0, st>>>(w, gamma, beta, mean, var, eps, w_out, b_out, N, K);
    EGR_LAUNCHED();
    return EGR_OK;
}

int stem_repack(const float* w, float* out, int N, cudaStream_t st) {
    stem_repack_kernel<<<grid_for((int64_t)N * STEM_K), 256, 0, st>>>(w, out, N);
    EGR_LAUNCHED();
    return EGR_OK;
}

}  // namespace egr
