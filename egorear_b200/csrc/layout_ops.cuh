// HBM-bound layout / resampling stages around the dense GEMMs (declarations; see layout_ops.cu).
#pragma once
#include "common.cuh"

namespace egr {

// in  [B][V][C][HW] fp32 (NCHW per view)  ->  out [V][B][HW][C] (view-major, channels-last)
// out_mode: 0 fp32, 1 bf16, 2 fp32 rounded to the nearest TF32 value (operand of a kind::tf32 stage), 3 fp16
int nchw_to_nhwc(const float* in, void* out, int B, int V, int C, int HW, int out_mode, cudaStream_t st);
// fp32 -> activation dtype copy (out_bf16: 0 float = plain copy, 1 bf16, 2 fp16)
int cast_act(const float* in, void* out, int out_bf16, int64_t n, cudaStream_t st);

// Heatmap head tail over G groups of B images:  z [g][B][Hs*Ws][C] (pre-activation, channels-last)
//   hm[b*hm_bs + g*hm_gs + (j*H + y)*W + x] fp32 = W15 · relu(up2(z)) + b15, group g uses weight set wsel[g]
//   optional copy hm_t [g][B][J][4HsWs] in the activation dtype (A operand of the jqa heatmap_proj GEMM)
int head_up_conv(const void* z, int z_bf16, const float* w, const float* bias, const int* wsel_host, int B, int G,
                 int Hs, int Ws, int C, int J, float* hm, int64_t hm_bs, int64_t hm_gs, void* hm_t,
                 cudaStream_t st);

// the same tail on the tensor cores: fp16 z (z_bf16 == 2 / 3 in head_up_conv), 32x32 -> 64x64, C = 128 (head_tail_tc.cu);
// hm_t bf16, or with precise != 0 (z_bf16 == 3, EGR_PREC_FP16) fp16 and the 1x1 weights as an fp16 hi + lo pair
int head_tail_tc(const void* z, const float* w, const float* bias, const int* wsel_host, int B, int G, int J, float* hm,
                 int64_t hm_bs, int64_t hm_gs, void* hm_t, int precise, cudaStream_t st);
// second generation (head_tail_mma.cu, option "tail_mma", default): the bilinear upsample as a tcgen05.mma against integer
// interpolation matrices (z read as an MN-major operand), fp32 row blend, then the 1x1 conv; persistent, TMA-fed
int head_tail_mma(const void* z, const float* w, const float* bias, const int* wsel_host, int B, int G, int J, float* hm,
                  int64_t hm_bs, int64_t hm_gs, void* hm_t, int precise, cudaStream_t st);
// R1 tail on the same kernel family: relu(up2(z)) -> ONE 16-bit channels-last map (z fp16 [n_img][32*32][128])
int up2_relu_mma(const void* z, int n_img, void* out, int out_f16, cudaStream_t st);
extern int g_opt_tail_mma;

// R1 tail: z [g][B][Hs*Ws][C] (z_dt: 0 fp32, 1 bf16, 2 fp16) -> relu(up2(z)) written to
//   out_nchw[b*o_bs + g*o_gs + (c*H + y)*W + x] fp32 (module output; optional) and up to two channels-last copies
//   cl0 / cl1 [g][B][4HsWs][C] with their own element types (cl*_dt: 0 fp32 rounded to TF32, 1 bf16, 2 fp16, 3 fp32 as is): the input of
//   the H2 3x3 conv and the high-precision copy of the pose3d proposal branch (one shared fp16 copy when chained)
int up2_relu_dual(const void* z, int z_dt, int B, int G, int Hs, int Ws, int C, float* out_nchw, int64_t o_bs,
                  int64_t o_gs, void* cl0, int cl0_dt, void* cl1, int cl1_dt, cudaStream_t st, int fp32_interp = 0,
                  bool allow_mma = true);
// fp32_interp: an fp16 z is interpolated in fp32 arithmetic (one rounding, EGR_PREC_FP16) instead of half2 (one per HFMA2)
// allow_mma: take the tensor-core kernel (up2_relu_mma) when only cl0 is wanted and option "tail_mma" is on

// nn.MaxPool2d(2) on channels-last [img][H][W][C] -> [img][H/2][W/2][C];  dt: 0 fp32, 1 bf16, 2 fp16
int maxpool2_nhwc(const void* in, void* out, int dt, int64_t n_img, int H, int W, int C, cudaStream_t st);

// ---- one-time weight preparation (prepack) ----
// conv weight [Cout][Cin][3][3] -> [Cout][ky][kx][Cin]
int repack_conv3(const float* w, float* out, int Cout, int Cin, cudaStream_t st);
// in-place fp32 -> nearest TF32 value
int round_tf32_inplace(float* p, int64_t n, cudaStream_t st);
// fp32 -> bf16 copy
int cast_bf16(const float* in, __nv_bfloat16* out, int64_t n, cudaStream_t st);
// fp32 -> fp16 copy (saturating)
int cast_f16(const float* in, __half* out, int64_t n, cudaStream_t st);
// split weights of the EGR_PREC_FP16 stages: in [R][K] fp32 -> out [R][2K] fp16 = [hi | lo], hi = fp16(w), lo = fp16(w - hi)
int split_f16(const float* in, __half* out, int64_t R, int K, cudaStream_t st);
// "3x TF32" weights of the fp32-grade token Linears: out [R][3K] fp32 = [hi | hi | lo], hi = tf32(w), lo = tf32(w - hi)
int split3_tf32(const float* in, float* out, int64_t R, int K, cudaStream_t st);
// fp16 flavour of the same product: [R][K] fp32 -> [R][3K] fp16 = [W_hi | W_hi | W_lo], A rows [x_hi | x_lo] fp16 (half the bytes)
int split3_f16(const float* in, __half* out, int64_t R, int K, cudaStream_t st);
// [R][C] -> [C][R]
int transpose2d(const float* in, float* out, int R, int C, cudaStream_t st);
// C[M][N] = A[M][K] · B[K][N] (+ bias[N] broadcast when non-null); small prepack-time products, fp32
int small_matmul(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, cudaStream_t st);
// out[r][c] = in[r][c] + v[c]
int add_rowvec(const float* in, const float* v, float* out, int64_t R, int C, cudaStream_t st);
// pose3d big-MLP weight: [N][(v c h w)] -> [N][(v h w c)]   (HWp = 64 positions, C = 128 channels)
int permute_mlp_weight(const float* w, float* out, int N, int V, int C, int HWp, cudaStream_t st);

}  // namespace egr
