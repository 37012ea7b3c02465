// Training-side operators for the modules of the hot path, so that `run.py fit` can stay on the
// extension for the pieces that have no PyTorch implementation (the mmcv op) or that dominate a step's small-kernel count:
//   egr_msda_backward   backward of mmcv's ms_deform_attn (single level): grad_value (atomic scatter), grad_sampling_loc,
//                       grad_attn_weight                                  (call site models/utils/deform_attn.py:155-162)
//   egr_mse_loss_*      nn.MSELoss(reduction="mean") forward + backward   (pl_wrappers/egoposeformer/heatmap_mvf_ex.py:258-261)
//   egr_mpjpe_loss_*    MpjpeLoss: mean over (batch, joint) of the L2 distance, forward + backward (models/utils/pose_metric.py:10-16)
// All fp32 (training runs at precision 32, configs/*:72); reductions are deterministic (fixed-order two-pass sums).
#include "token_common.cuh"

namespace egr {
namespace {

// one warp per (b, q, head); lanes stride the head's channels.  For every point: the 4 corner weights, then
//   grad_value[corner][c] += aw * w_corner * go[c]                                   (atomicAdd: several queries hit a pixel)
//   grad_aw[p]  = sum_c go[c] * bilinear(value)[c]
//   grad_loc[p] = (W * sum_c go[c] * aw * d bilinear / d w_im ,  H * sum_c go[c] * aw * d bilinear / d h_im)
// with out-of-map corners contributing 0 to both the value and its derivative (mmcv's ms_deform_attn_col2im_bilinear).
__global__ void __launch_bounds__(256)
msda_backward_kernel(const float* __restrict__ value, int B, int H, int W, int nh, int hd, const float* __restrict__ loc,
                     const float* __restrict__ aw, int Q, int P, const float* __restrict__ grad_out,
                     float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_aw) {
    const int lane = threadIdx.x & 31;
    const int64_t unit = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t total = (int64_t)B * Q * nh;
    if (unit >= total) return;
    const int h = (int)(unit % nh);
    const int64_t bq = unit / nh;
    const int b = (int)(bq / Q);
    const float* lp = loc + unit * P * 2;
    const float* ap = aw + unit * P;
    const float* go = grad_out + bq * nh * hd + (int64_t)h * hd;
    const int64_t vstride = (int64_t)nh * hd;
    const float* vb = value + (int64_t)b * H * W * vstride + (int64_t)h * hd;
    float* gvb = grad_value + (int64_t)b * H * W * vstride + (int64_t)h * hd;
    for (int p = 0; p < P; ++p) {
        const float lx = __ldg(lp + p * 2), ly = __ldg(lp + p * 2 + 1), a = __ldg(ap + p);
        const float h_im = ly * (float)H - 0.5f, w_im = lx * (float)W - 0.5f;
        float g_a = 0.f, g_w = 0.f, g_h = 0.f;
        if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
            const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
            const int h_high = h_low + 1, w_high = w_low + 1;
            const float lh = h_im - (float)h_low, lw = w_im - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
            const bool y0 = h_low >= 0, y1 = h_high <= H - 1, x0 = w_low >= 0, x1 = w_high <= W - 1;
            for (int c = lane; c < hd; c += 32) {
                const float g = go[c];
                const float v00 = (y0 && x0) ? __ldg(vb + ((int64_t)h_low * W + w_low) * vstride + c) : 0.f;
                const float v01 = (y0 && x1) ? __ldg(vb + ((int64_t)h_low * W + w_high) * vstride + c) : 0.f;
                const float v10 = (y1 && x0) ? __ldg(vb + ((int64_t)h_high * W + w_low) * vstride + c) : 0.f;
                const float v11 = (y1 && x1) ? __ldg(vb + ((int64_t)h_high * W + w_high) * vstride + c) : 0.f;
                const float ga = g * a;
                if (y0 && x0) atomicAdd(gvb + ((int64_t)h_low * W + w_low) * vstride + c, ga * hh * hw);
                if (y0 && x1) atomicAdd(gvb + ((int64_t)h_low * W + w_high) * vstride + c, ga * hh * lw);
                if (y1 && x0) atomicAdd(gvb + ((int64_t)h_high * W + w_low) * vstride + c, ga * lh * hw);
                if (y1 && x1) atomicAdd(gvb + ((int64_t)h_high * W + w_high) * vstride + c, ga * lh * lw);
                g_a += g * (hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11);
                g_w += ga * (hh * (v01 - v00) + lh * (v11 - v10));
                g_h += ga * (hw * (v10 - v00) + lw * (v11 - v01));
            }
        }
        g_a = warp_sum(g_a); g_w = warp_sum(g_w); g_h = warp_sum(g_h);
        if (lane == 0) {
            grad_aw[unit * P + p] = g_a;
            grad_loc[(unit * P + p) * 2 + 0] = g_w * (float)W;
            grad_loc[(unit * P + p) * 2 + 1] = g_h * (float)H;
        }
    }
}

constexpr int RED_BLOCKS = 592;      // 4 x 148 CTAs; partial sums are combined in a fixed order

// partial[blockIdx] = sum over this block's grid-stride slice of f(a[i], b[i]);  MODE 0: squared difference
__global__ void __launch_bounds__(256)
sqdiff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        s += (double)(d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}
// partial sums of the per-joint L2 distances; one thread per joint (D coordinates each)
__global__ void __launch_bounds__(256)
l2_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n_joints, int D, double* __restrict__ partial) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_joints; j += (int64_t)gridDim.x * blockDim.x) {
        float q = 0.f;
        for (int d = 0; d < D; ++d) { const float x = b[j * D + d] - a[j * D + d]; q = fmaf(x, x, q); }
        s += (double)sqrtf(q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void finish_mean_kernel(const double* __restrict__ partial, int n_partial, double inv_n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < n_partial; ++i) t += partial[i];
        out[0] = (float)(t * inv_n);
    }
}
// d mean((a-b)^2) / da = 2 (a - b) / n * g
__global__ void __launch_bounds__(256)
mse_backward_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g, int64_t n, float scale,
                    float* __restrict__ ga) {
    const float gs = g[0] * scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        ga[i] = (a[i] - b[i]) * gs;
}
// d mean_j ||gt_j - pred_j|| / d pred_j = (pred_j - gt_j) / (||.|| * n_joints) * g   (0 where the distance is 0, like torch's norm backward)
__global__ void __launch_bounds__(256)
mpjpe_backward_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ g, int64_t n_joints,
                      int D, float* __restrict__ gp) {
    const float gs = g[0] / (float)n_joints;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_joints; j += (int64_t)gridDim.x * blockDim.x) {
        float q = 0.f;
        for (int d = 0; d < D; ++d) { const float x = pred[j * D + d] - gt[j * D + d]; q = fmaf(x, x, q); }
        const float inv = q > 0.f ? gs / sqrtf(q) : 0.f;
        for (int d = 0; d < D; ++d) gp[j * D + d] = (pred[j * D + d] - gt[j * D + d]) * inv;
    }
}

inline int grid_for(int64_t n) {
    const int64_t g = ceil_div64(n, 256);
    return (int)(g < RED_BLOCKS ? (g > 0 ? g : 1) : RED_BLOCKS);
}

}  // namespace
}  // namespace egr

using namespace egr;

extern "C" int egr_msda_backward(const float* value, int B, int H, int W, int nh, int hd, const float* loc, const float* aw,
                                 int Q, int P, const float* grad_out, float* grad_value, float* grad_loc, float* grad_aw,
                                 void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(B >= 0 && H > 0 && W > 0 && nh > 0 && hd > 0 && Q > 0 && P > 0, EGR_ERR_INVALID, "msda_backward: bad shape");
    if (B == 0) return EGR_OK;
    EGR_CHECK(value && loc && aw && grad_out && grad_value && grad_loc && grad_aw, EGR_ERR_INVALID, "msda_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    EGR_CUDA_OK(cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)B * H * W * nh * hd, st));
    const int64_t units = (int64_t)B * Q * nh;
    msda_backward_kernel<<<(int)ceil_div64(units, 8), 256, 0, st>>>(value, B, H, W, nh, hd, loc, aw, Q, P, grad_out, grad_value,
                                                                      grad_loc, grad_aw);
    EGR_LAUNCHED();
    return EGR_OK;
}

extern "C" int64_t egr_loss_workspace_bytes(void) { return (int64_t)sizeof(double) * RED_BLOCKS; }

extern "C" int egr_mse_loss_forward(const float* pred, const float* target, int64_t n, float* loss, void* workspace, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(n > 0 && pred && target && loss && workspace, EGR_ERR_INVALID, "mse_loss: null pointer / empty input");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n);
    sqdiff_partial_kernel<<<g, 256, 0, st>>>(pred, target, n, (double*)workspace);
    EGR_LAUNCHED();
    finish_mean_kernel<<<1, 32, 0, st>>>((const double*)workspace, g, 1.0 / (double)n, loss);
    EGR_LAUNCHED();
    return EGR_OK;
}

extern "C" int egr_mse_loss_backward(const float* pred, const float* target, const float* grad_loss, int64_t n, float* grad_pred,
                                     void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(n > 0 && pred && target && grad_loss && grad_pred, EGR_ERR_INVALID, "mse_loss_backward: null pointer / empty input");
    mse_backward_kernel<<<grid_for(n) * 4, 256, 0, (cudaStream_t)stream>>>(pred, target, grad_loss, n, 2.0f / (float)n, grad_pred);
    EGR_LAUNCHED();
    return EGR_OK;
}

extern "C" int egr_mpjpe_loss_forward(const float* pred, const float* gt, int64_t n_joints, int D, float* loss, void* workspace,
                                      void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(n_joints > 0 && D > 0 && pred && gt && loss && workspace, EGR_ERR_INVALID, "mpjpe_loss: null pointer / empty input");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n_joints);
    l2_partial_kernel<<<g, 256, 0, st>>>(pred, gt, n_joints, D, (double*)workspace);
    EGR_LAUNCHED();
    finish_mean_kernel<<<1, 32, 0, st>>>((const double*)workspace, g, 1.0 / (double)n_joints, loss);
    EGR_LAUNCHED();
    return EGR_OK;
}

extern "C" int egr_mpjpe_loss_backward(const float* pred, const float* gt, const float* grad_loss, int64_t n_joints, int D,
                                       float* grad_pred, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(n_joints > 0 && D > 0 && pred && gt && grad_loss && grad_pred, EGR_ERR_INVALID, "mpjpe_loss_backward: null pointer");
    mpjpe_backward_kernel<<<grid_for(n_joints), 256, 0, (cudaStream_t)stream>>>(pred, gt, grad_loss, n_joints, D, grad_pred);
    EGR_LAUNCHED();
    return EGR_OK;
}
