// Backbone-side kernels around the tcgen05 conv stages (declarations; see backbone_ops.cu).
#pragma once
#include "common.cuh"

namespace egr {

constexpr int STEM_K = 192;     // 7*7*3 = 147 taps x channels, zero padded to a multiple of the 64-element k-block

// img [B][V][3][H][W] fp32 -> out [V][B][(H/2)*(W/2)][STEM_K] 16-bit (dt: 1 bf16, 2 fp16): rows of the 7x7 stride-2 pad-3 stem conv
int stem_im2col(const float* img, void* out, int dt, int B, int V, int H, int W, cudaStream_t st);
// the stem as ONE tensor-core kernel without the im2col buffer (stem_tc.cu): img [B][V][3][256][256] fp32, w16 [sets][64][STEM_K]
// 16-bit (BN folded, im2col order), bias [sets][64] -> out [V][B][128*128][64] 16-bit = relu(conv + bias); views [0, vpg) use set 0
int stem_tc(const float* img, const void* w16, const float* bias, int dt, int B, int V, int vpg, void* out, cudaStream_t st);
// nn.MaxPool2d(3, stride 2, padding 1) on channels-last [img][H][W][C] 16-bit -> [img][H/2][W/2][C]
int maxpool3s2_nhwc(const void* in, void* out, int dt, int64_t n_img, int H, int W, int C, cudaStream_t st);
// EfficientFPN top-down step after the two 1x1 halves of the fuse conv: out = relu(hi + bilinear_x2_align_corners(low)),
// hi / out [img][R*R][C], low [img][(R/2)^2][C], 16-bit channels-last; out may alias hi
int up2_add_relu_nhwc(const void* hi, const void* low, void* out, int dt, int64_t n_img, int R, int C, cudaStream_t st);
// in [V][B][HW][C] 16-bit -> out [B][V][C][HW] fp32
int bottom_to_nchw(const void* in, float* out, int dt, int B, int V, int HW, int C, cudaStream_t st);
// eval-mode BatchNorm folded into the conv in front of it: w_out[n][k] = w[n][k] * g[n] / sqrt(var[n] + eps), b_out[n] = beta[n] - mean[n] * (..)
int bn_fold(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* w_out,
            float* b_out, int N, int K, cudaStream_t st);
// stem weight [N][3][7][7] -> [N][STEM_K] in im2col order (ky, kx, c), zero padded
int stem_repack(const float* w, float* out, int N, cudaStream_t st);

}  // namespace egr
