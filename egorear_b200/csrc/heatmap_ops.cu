// G1 generate_target, D1 get_max_preds, H1' 1x1 heatmap head, joint packing.
// All four are HBM-bound streaming kernels: one warp owns one heatmap, 128-bit accesses,
// 512 contiguous bytes per warp instruction, streaming cache hints (data is touched once).
#include "common.cuh"

namespace egr {

// ---------------------------------------------------------------------------------------------
// G1  generate_heatmap.py:10-48
//   mu = int(coord / feat_stride + 0.5) in float64 with C truncation, ul = mu - 3s, br = mu + 3s + 1,
//   all-zero map when the patch misses the map entirely, else the clipped fp32 patch is pasted.
// ---------------------------------------------------------------------------------------------
constexpr int GT_WARPS = 8;
constexpr int GT_MAX_PATCH = 31;
// The Gaussian patch travels BY VALUE as a kernel parameter (3.8 KB, constant bank): no device-side cache that another
// stream / device / sigma could overwrite while a launch is still reading it, no host->device copy (legal under stream
// capture), nothing to free.
struct GtPatch { float v[GT_MAX_PATCH * GT_MAX_PATCH]; };

__global__ void __launch_bounds__(GT_WARPS * 32)
generate_target_kernel(const double* __restrict__ joints, float* __restrict__ out, int64_t n_hm,
                       double feat_stride, double tmp_size, int hs, int size, const __grid_constant__ GtPatch patch) {
    __shared__ float s_patch[GT_MAX_PATCH * GT_MAX_PATCH];
    for (int i = threadIdx.x; i < size * size; i += blockDim.x) s_patch[i] = patch.v[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int row_v4 = hs >> 2;              // float4 per row
    const int n_v4 = hs * row_v4;            // float4 per heatmap
    for (int64_t hm = (int64_t)blockIdx.x * GT_WARPS + warp; hm < n_hm; hm += (int64_t)gridDim.x * GT_WARPS) {
        // every lane evaluates the (cheap, uniform) index math; the two loads are a warp broadcast
        const double jx = joints[hm * 2 + 0];
        const double jy = joints[hm * 2 + 1];
        const int mu_x = (int)(jx / feat_stride + 0.5);   // IEEE double divide + add, then truncation
        const int mu_y = (int)(jy / feat_stride + 0.5);
        const int ul0 = (int)(mu_x - tmp_size), ul1 = (int)(mu_y - tmp_size);
        const int br0 = (int)(mu_x + tmp_size + 1.0), br1 = (int)(mu_y + tmp_size + 1.0);
        const bool skip = (ul0 >= hs) || (ul1 >= hs) || (br0 < 0) || (br1 < 0);
        // image-space window that receives patch values: [ix0, ix1) x [iy0, iy1)
        const int ix0 = max(0, ul0), ix1 = skip ? 0 : min(br0, hs);
        const int iy0 = max(0, ul1), iy1 = skip ? 0 : min(br1, hs);
        float4* dst = reinterpret_cast<float4*>(out + hm * (int64_t)hs * hs);
        for (int i = lane; i < n_v4; i += 32) {
            const int y = i / row_v4;
            const int x = (i - y * row_v4) << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= iy0 && y < iy1 && x + 3 >= ix0 && x < ix1) {
                const int pb = (y - ul1) * size - ul0 + x;              // patch[(y-ul1)*size + (x-ul0)]
                if (x + 0 >= ix0 && x + 0 < ix1) v.x = s_patch[pb + 0];
                if (x + 1 >= ix0 && x + 1 < ix1) v.y = s_patch[pb + 1];
                if (x + 2 >= ix0 && x + 2 < ix1) v.z = s_patch[pb + 2];
                if (x + 3 >= ix0 && x + 3 < ix1) v.w = s_patch[pb + 3];
            }
            __stcs(dst + i, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// D1  utils/loss.py:122-142 — one warp per heatmap, first index wins, NaN propagates (torch.max)
// ---------------------------------------------------------------------------------------------
constexpr int DEC_WARPS = 8;

__device__ __forceinline__ bool better(float v, int i, float best, int bi) {
    // ordering: NaN > everything; among equals the smaller index wins
    const bool vn = (v != v), bn = (best != best);
    if (vn || bn) return vn && (!bn || i < bi);
    return (v > best) || (v == best && i < bi);
}

__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_argmax_kernel(const float* __restrict__ hm, int64_t n_hm, int HW, int W, float inv_w, float inv_h,
                     int H, float threshold, int normalize, float* __restrict__ preds,
                     float* __restrict__ maxvals, uint8_t* __restrict__ valid, int32_t* __restrict__ idx_out) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_v4 = HW >> 2;
    for (int64_t m = (int64_t)blockIdx.x * DEC_WARPS + warp; m < n_hm; m += (int64_t)gridDim.x * DEC_WARPS) {
        const float4* src = reinterpret_cast<const float4*>(hm + m * (int64_t)HW);
        float best = -INFINITY;
        int bi = 0x7fffffff;
        bool seen_nan = false;
        int i = lane;
        // 4 independent 128-bit loads in flight per lane
        for (; i + 96 < n_v4; i += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldcs(src + i + u * 32);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int base = (i + u * 32) << 2;
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float x = e[k];
                    if (!seen_nan) {
                        if (x != x) { best = x; bi = base + k; seen_nan = true; }
                        else if (x > best || bi == 0x7fffffff) { best = x; bi = base + k; }
                    }
                }
            }
        }
        for (; i < n_v4; i += 32) {
            const float4 v = __ldcs(src + i);
            const int base = i << 2;
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float x = e[k];
                if (!seen_nan) {
                    if (x != x) { best = x; bi = base + k; seen_nan = true; }
                    else if (x > best || bi == 0x7fffffff) { best = x; bi = base + k; }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ob, oi, best, bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
            float x = (float)(bi % W), y = (float)(bi / W);
            if (normalize) { x = x / (float)W; y = y / (float)H; }
            preds[m * 2 + 0] = x;
            preds[m * 2 + 1] = y;
            maxvals[m] = best;
            valid[m] = (uint8_t)(best >= threshold);
            if (idx_out) idx_out[m] = bi;
        }
    }
    (void)inv_w; (void)inv_h;
}

// ---------------------------------------------------------------------------------------------
// D1s  soft-argmax decoder: get_max_preds_soft_pytorch (pose_estimation/utils/loss.py:145-177)
//   p = softmax over the H*W values of a map; x = sum_w w * sum_h p[h][w], y = sum_h h * sum_w p[h][w]; maxvals = max.
//   One warp per map, two passes (the 16 KB map of the second pass comes from L1/L2): pass 1 max (128-bit loads, shuffle
//   reduction), pass 2 the three sums sum e, sum e*w, sum e*h with e = exp(v - max) in fp32, one division at the end.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_soft_argmax_kernel(const float* __restrict__ hm, int64_t n_hm, int H, int W, int normalize,
                          float* __restrict__ preds, float* __restrict__ maxvals) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int HW = H * W, n_v4 = HW >> 2;
    for (int64_t m = (int64_t)blockIdx.x * DEC_WARPS + warp; m < n_hm; m += (int64_t)gridDim.x * DEC_WARPS) {
        const float4* src = reinterpret_cast<const float4*>(hm + m * (int64_t)HW);
        float mx = -INFINITY;
        bool nan = false;      // fmaxf drops NaN; torch.max / softmax propagate it (maxvals and both coordinates become NaN)
        for (int i = lane; i < n_v4; i += 32) {
            const float4 v = __ldg(src + i);
            mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            nan = nan || (v.x != v.x) || (v.y != v.y) || (v.z != v.z) || (v.w != v.w);
        }
        mx = warp_max(mx);
        if (__any_sync(0xffffffffu, nan)) mx = __int_as_float(0x7fc00000);
        float s = 0.f, sx = 0.f, sy = 0.f;
        for (int i = lane; i < n_v4; i += 32) {
            const float4 v = __ldg(src + i);
            const int base = i << 2;                 // W % 4 == 0: the 4 values share a row
            const int h = base / W, w0 = base - h * W;
            const float e0 = expf(v.x - mx), e1 = expf(v.y - mx), e2 = expf(v.z - mx), e3 = expf(v.w - mx);
            const float es = (e0 + e1) + (e2 + e3);
            s += es;
            sy = fmaf(es, (float)h, sy);
            sx += e0 * (float)w0 + e1 * (float)(w0 + 1) + e2 * (float)(w0 + 2) + e3 * (float)(w0 + 3);
        }
        s = warp_sum(s); sx = warp_sum(sx); sy = warp_sum(sy);
        if (lane == 0) {
            float x = sx / s, y = sy / s;
            if (normalize) { x = x / (float)W; y = y / (float)H; }
            preds[m * 2 + 0] = x;
            preds[m * 2 + 1] = y;
            maxvals[m] = mx;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// D1i  integral (soft-argmax) decoder: integrate_tensor_2d (pose_estimation/utils/util.py:80-109)
//   v = hm * multiplier; softmax: p = softmax(v) over H*W, else p = relu(v); x = sum_w w * sum_h p, y = sum_h h * sum_w p,
//   divided by the total mass in the relu variant (0/0 -> NaN like torch); returns the coordinates AND p.
//   One warp per map: max pass, sums pass, write pass (passes 2 and 3 re-read the 16 KB map from L1/L2): HBM traffic is
//   one read + one write of the map.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_WARPS * 32)
integrate_2d_kernel(const float* __restrict__ hm, int64_t n_hm, int H, int W, int softmax, float mult,
                    float* __restrict__ coords, float* __restrict__ hm_out) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int HW = H * W, n_v4 = HW >> 2;
    for (int64_t m = (int64_t)blockIdx.x * DEC_WARPS + warp; m < n_hm; m += (int64_t)gridDim.x * DEC_WARPS) {
        const float4* src = reinterpret_cast<const float4*>(hm + m * (int64_t)HW);
        float4* dst = hm_out ? reinterpret_cast<float4*>(hm_out + m * (int64_t)HW) : nullptr;
        float mx = 0.f;
        if (softmax) {
            mx = -INFINITY;
            bool nan = false;  // a NaN anywhere in the map makes the softmax (all of p and both coordinates) NaN, as in torch
            for (int i = lane; i < n_v4; i += 32) {
                const float4 v = __ldg(src + i);
                mx = fmaxf(fmaxf(mx, fmaxf(v.x * mult, v.y * mult)), fmaxf(v.z * mult, v.w * mult));
                nan = nan || (v.x != v.x) || (v.y != v.y) || (v.z != v.z) || (v.w != v.w);
            }
            mx = warp_max(mx);
            if (__any_sync(0xffffffffu, nan)) mx = __int_as_float(0x7fc00000);
        }
        float s = 0.f, sx = 0.f, sy = 0.f;
        for (int i = lane; i < n_v4; i += 32) {
            const float4 v = __ldg(src + i);
            const int base = i << 2;                 // W % 4 == 0: the 4 values share a row
            const int h = base / W, w0 = base - h * W;
            float e0, e1, e2, e3;
            if (softmax) {
                // hm * multiplier is rounded BEFORE the softmax in the reference: no FMA contraction into (v*mult - mx)
                e0 = expf(__fmul_rn(v.x, mult) - mx); e1 = expf(__fmul_rn(v.y, mult) - mx);
                e2 = expf(__fmul_rn(v.z, mult) - mx); e3 = expf(__fmul_rn(v.w, mult) - mx);
            } else {
                // torch.relu keeps NaN (fmaxf would drop it)
                const float m0 = v.x * mult, m1 = v.y * mult, m2 = v.z * mult, m3 = v.w * mult;
                e0 = (m0 != m0) ? m0 : fmaxf(m0, 0.f); e1 = (m1 != m1) ? m1 : fmaxf(m1, 0.f);
                e2 = (m2 != m2) ? m2 : fmaxf(m2, 0.f); e3 = (m3 != m3) ? m3 : fmaxf(m3, 0.f);
                if (dst) __stcs(dst + i, make_float4(e0, e1, e2, e3));
            }
            const float es = (e0 + e1) + (e2 + e3);
            s += es;
            sy = fmaf(es, (float)h, sy);
            sx += e0 * (float)w0 + e1 * (float)(w0 + 1) + e2 * (float)(w0 + 2) + e3 * (float)(w0 + 3);
        }
        s = warp_sum(s); sx = warp_sum(sx); sy = warp_sum(sy);
        if (softmax && dst) {
            for (int i = lane; i < n_v4; i += 32) {
                const float4 v = __ldg(src + i);
                __stcs(dst + i, make_float4(expf(__fmul_rn(v.x, mult) - mx) / s, expf(__fmul_rn(v.y, mult) - mx) / s,
                                            expf(__fmul_rn(v.z, mult) - mx) / s, expf(__fmul_rn(v.w, mult) - mx) / s));
            }
        }
        if (lane == 0) {
            coords[m * 2 + 0] = sx / s;
            coords[m * 2 + 1] = sy / s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// H1'  1x1 conv C->J on NCHW fp32 (estimator/egoposeformer_heatmap.py:23,34-39)
//   thread = 4 consecutive positions; loads are float4-coalesced per channel; weights broadcast from smem
// ---------------------------------------------------------------------------------------------
constexpr int H1_JMAX = 16;

__global__ void __launch_bounds__(256)
heatmap_head_1x1_kernel(const float* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ b,
                        int64_t n_img, int C, int HW, int J, float* __restrict__ out) {
    extern __shared__ float s_w[];   // [C][H1_JMAX]
    for (int i = threadIdx.x; i < C * H1_JMAX; i += blockDim.x) {
        const int c = i / H1_JMAX, j = i % H1_JMAX;
        s_w[i] = (j < J) ? w[j * C + c] : 0.f;
    }
    __syncthreads();
    const int v4_per_img = HW >> 2;
    const int64_t total = n_img * v4_per_img;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = t / v4_per_img;
        const int p4 = (int)(t - n * v4_per_img);
        const float4* src = reinterpret_cast<const float4*>(feat + n * (int64_t)C * HW) + p4;
        float4 acc[H1_JMAX];
#pragma unroll
        for (int j = 0; j < H1_JMAX; ++j) {
            const float bj = (j < J) ? b[j] : 0.f;
            acc[j] = make_float4(bj, bj, bj, bj);
        }
        for (int c = 0; c < C; ++c) {
            const float4 x = __ldcs(src + (int64_t)c * v4_per_img);
            const float4* wr = reinterpret_cast<const float4*>(s_w + c * H1_JMAX);
#pragma unroll
            for (int q = 0; q < H1_JMAX / 4; ++q) {
                const float4 ww = wr[q];
                const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4& a = acc[q * 4 + k];
                    a.x = fmaf(wv[k], x.x, a.x); a.y = fmaf(wv[k], x.y, a.y);
                    a.z = fmaf(wv[k], x.z, a.z); a.w = fmaf(wv[k], x.w, a.w);
                }
            }
        }
        float4* dst = reinterpret_cast<float4*>(out + n * (int64_t)J * HW) + p4;
#pragma unroll
        for (int j = 0; j < H1_JMAX; ++j)
            if (j < J) __stcs(dst + (int64_t)j * v4_per_img, acc[j]);
    }
}

__global__ void pack_joints_kernel(const float* __restrict__ p2, const float* __restrict__ p3, int B, int n2, int n3,
                                   float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int row = n2 + n3;
    const int64_t total = (int64_t)B * row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / row;
        const int k = (int)(i - b * row);
        out[i] = (k < n2) ? p2[b * n2 + k] : p3[b * n3 + (k - n2)];
    }
}

}  // namespace egr

using namespace egr;

// recorded numpy fp32 values of exp(-d2/2), d2 = dx^2+dy^2 (checked in tests)
static float sigma1_value(int d2) {
    uint32_t b = 0;
    switch (d2) {
        case 0: b = 0x3F800000u; break;  case 1: b = 0x3F1B4598u; break;
        case 2: b = 0x3EBC5AB1u; break;  case 4: b = 0x3E0A9555u; break;
        case 5: b = 0x3DA81C2Fu; break;  case 8: b = 0x3C960AAEu; break;
        case 9: b = 0x3C360282u; break;  case 10: b = 0x3BDCC9FEu; break;
        case 13: b = 0x3AC50F0Cu; break; case 18: b = 0x39016791u; break;
    }
    float v;
    memcpy(&v, &b, 4);
    return v;
}


extern "C" int egr_generate_target(const double* joints, float* out, int64_t n_maps, int J, double image_size,
                                   int heatmap_size, double sigma, const float* patch_host, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(n_maps >= 0 && J > 0, EGR_ERR_INVALID, "generate_target: n_maps=%lld J=%d", (long long)n_maps, J);
    EGR_CHECK(heatmap_size > 0 && heatmap_size % 4 == 0, EGR_ERR_UNSUPPORTED,
              "generate_target: heatmap_size must be a positive multiple of 4, got %d", heatmap_size);
    EGR_CHECK(sigma > 0 && image_size > 0, EGR_ERR_INVALID, "generate_target: sigma and image_size must be > 0");
    const double tmp_size = sigma * 3.0;
    EGR_CHECK(tmp_size == (double)(int)tmp_size, EGR_ERR_UNSUPPORTED,
              "generate_target: 3*sigma must be an integer (reference slicing is undefined otherwise), sigma=%g", sigma);
    const int size = 2 * (int)tmp_size + 1;
    EGR_CHECK(size <= GT_MAX_PATCH, EGR_ERR_UNSUPPORTED, "generate_target: patch %d > %d", size, GT_MAX_PATCH);
    if (n_maps == 0) return EGR_OK;
    EGR_CHECK(joints && out, EGR_ERR_INVALID, "generate_target: null pointer");
    EGR_CHECK(((uintptr_t)out % 16) == 0 && ((uintptr_t)joints % 8) == 0, EGR_ERR_INVALID,
              "generate_target: out must be 16-byte aligned (128-bit stores), joints 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;

    GtPatch h;
    memset(&h, 0, sizeof(h));
    if (patch_host) {
        memcpy(h.v, patch_host, sizeof(float) * size * size);
    } else {
        const int c = size / 2;
        for (int y = 0; y < size; ++y)
            for (int x = 0; x < size; ++x) {
                const int d2 = (x - c) * (x - c) + (y - c) * (y - c);
                h.v[y * size + x] = (sigma == 1.0) ? sigma1_value(d2)
                                                   : expf(-(float)d2 / (float)(2.0 * sigma * sigma));
            }
    }
    const int64_t n_hm = n_maps * J;
    const int64_t want = ceil_div64(n_hm, GT_WARPS);
    const int grid = (int)(want < (int64_t)sm_count() * 32 ? want : (int64_t)sm_count() * 32);
    generate_target_kernel<<<grid, GT_WARPS * 32, 0, st>>>(joints, out, n_hm, image_size / (double)heatmap_size,
                                                            tmp_size, heatmap_size, size, h);
    EGR_LAUNCHED();
    return EGR_OK;
}

extern "C" int egr_decode_argmax(const float* hm, int64_t N, int J, int H, int W, float threshold, int normalize,
                                 float* preds, float* maxvals, uint8_t* valid, int32_t* idx, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(N >= 0 && J > 0 && H > 0 && W > 0, EGR_ERR_INVALID, "decode: heatmaps should be 4-ndim [N,J,H,W]");
    EGR_CHECK((H * W) % 4 == 0, EGR_ERR_UNSUPPORTED, "decode: H*W must be a multiple of 4");
    if (N == 0) return EGR_OK;
    EGR_CHECK(hm && preds && maxvals && valid, EGR_ERR_INVALID, "decode: null pointer");
    EGR_CHECK(((uintptr_t)hm % 16) == 0, EGR_ERR_INVALID, "decode: hm must be 16-byte aligned (128-bit loads)");
    const int64_t n_hm = N * J;
    const int64_t want = ceil_div64(n_hm, DEC_WARPS);
    const int grid = (int)(want < (int64_t)sm_count() * 32 ? want : (int64_t)sm_count() * 32);
    EGR_LAUNCH(decode_argmax_kernel, grid, DEC_WARPS * 32, 0, (cudaStream_t)stream,
               hm, n_hm, H * W, W, 1.f / W, 1.f / H, H, threshold, normalize, preds, maxvals, valid, idx);
    return EGR_OK;
}

extern "C" int egr_integrate_tensor_2d(const float* hm, int64_t N, int J, int H, int W, int softmax, float multiplier,
                                       float* coords, float* hm_out, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(N >= 0 && J > 0 && H > 0 && W > 0, EGR_ERR_INVALID, "integrate_tensor_2d: heatmaps should be [B, J, H, W]");
    EGR_CHECK(W % 4 == 0, EGR_ERR_UNSUPPORTED, "integrate_tensor_2d: W must be a multiple of 4");
    if (N == 0) return EGR_OK;
    EGR_CHECK(hm && coords, EGR_ERR_INVALID, "integrate_tensor_2d: null pointer");
    EGR_CHECK(((uintptr_t)hm % 16) == 0 && ((uintptr_t)hm_out % 16) == 0, EGR_ERR_INVALID, "integrate_tensor_2d: pointers must be 16-byte aligned");
    const int64_t n_hm = N * J;
    const int grid = (int)std::min<int64_t>(ceil_div64(n_hm, DEC_WARPS), (int64_t)sm_count() * 16);
    EGR_LAUNCH(integrate_2d_kernel, grid, DEC_WARPS * 32, 0, (cudaStream_t)stream, hm, n_hm, H, W, softmax, multiplier, coords, hm_out);
    return EGR_OK;
}

extern "C" int egr_decode_soft_argmax(const float* hm, int64_t N, int J, int H, int W, int normalize, float* preds,
                                      float* maxvals, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(N >= 0 && J > 0 && H > 0 && W > 0, EGR_ERR_INVALID, "decode_soft: batch_images should be 4-ndim (B, J, H, W)");
    EGR_CHECK(W % 4 == 0, EGR_ERR_UNSUPPORTED, "decode_soft: W must be a multiple of 4");
    if (N == 0) return EGR_OK;
    EGR_CHECK(hm && preds && maxvals, EGR_ERR_INVALID, "decode_soft: null pointer");
    EGR_CHECK(((uintptr_t)hm % 16) == 0, EGR_ERR_INVALID, "decode_soft: hm must be 16-byte aligned (128-bit loads)");
    const int64_t n_hm = N * J;
    const int64_t want = ceil_div64(n_hm, DEC_WARPS);
    const int grid = (int)(want < (int64_t)sm_count() * 32 ? want : (int64_t)sm_count() * 32);
    EGR_LAUNCH(decode_soft_argmax_kernel, grid, DEC_WARPS * 32, 0, (cudaStream_t)stream, hm, n_hm, H, W, normalize, preds, maxvals);
    return EGR_OK;
}

extern "C" int egr_heatmap_head_1x1(const float* feat, const float* weight, const float* bias, int64_t N, int C,
                                    int HW, int J, float* out, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(J > 0 && J <= H1_JMAX, EGR_ERR_UNSUPPORTED, "heatmap_head_1x1: J=%d > %d", J, H1_JMAX);
    EGR_CHECK(HW % 4 == 0 && C > 0, EGR_ERR_UNSUPPORTED, "heatmap_head_1x1: HW %% 4 != 0");
    if (N == 0) return EGR_OK;
    EGR_CHECK(feat && weight && bias && out, EGR_ERR_INVALID, "heatmap_head_1x1: null pointer");
    EGR_CHECK(((uintptr_t)feat % 16) == 0 && ((uintptr_t)out % 16) == 0, EGR_ERR_INVALID,
              "heatmap_head_1x1: feat / out must be 16-byte aligned (128-bit accesses)");
    const int64_t total = N * (HW / 4);
    const int64_t want = ceil_div64(total, 256);
    const int grid = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    const size_t smem = sizeof(float) * C * H1_JMAX;
    EGR_CHECK(smem <= 48 * 1024, EGR_ERR_UNSUPPORTED, "heatmap_head_1x1: C=%d too large", C);
    heatmap_head_1x1_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(feat, weight, bias, N, C, HW, J, out);
    EGR_LAUNCHED();
    return EGR_OK;
}

extern "C" int egr_pack_joints(const float* preds2d, const float* pose3d, int B, int n2d, int n3d, float* packed,
                               void* stream) {
    if (int rc = require_device()) return rc;
    if (B == 0) return EGR_OK;
    const int64_t total = (int64_t)B * (n2d + n3d);
    EGR_LAUNCH(pack_joints_kernel, (int)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream, preds2d, pose3d, B, n2d, n3d, packed);
    return EGR_OK;
}
