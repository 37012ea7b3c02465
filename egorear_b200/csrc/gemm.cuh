// Dense stage descriptor shared by the fp32 SIMT GEMM (gemm_simt.cu) and the tcgen05 bf16 GEMM (gemm_tc.cu).
//
// Every dense op on the hot path (1x1 conv, 3x3 stride-2 conv, Linear) is D = epi(A · W^T + bias) with
//   A  [M, K]  activations, channels-last (NHWC), K contiguous           (implicit im2col for the 3x3)
//   W  [N, K]  weights, K contiguous (PyTorch Linear / repacked conv weight [Cout][ky][kx][Cin])
//   D  [M, N]  channels-last output
// and `groups` independent problems (different weights: front/back heads, the 4 per-view refiners)
// launched together through blockIdx.z with per-group element strides.
#pragma once
#include "common.cuh"

namespace egr {

enum AMode : int {
    A_PLAIN = 0,      // A[m][k] = A[(k / kblk) * kblk_stride + m * lda + (k % kblk)]
    A_CONV3S2 = 1,    // 3x3, stride 2, pad 1 over NHWC [img][Hin][Win][Cin]; k = (ky*3+kx)*Cin + ci
                      // (K == Cin: only the centre tap = a 1x1 stride-2 conv, ResNet's downsample branch)
    A_CONV3S1 = 2,    // 3x3, stride 1, pad 1 over NHWC [img][Hin][Win][Cin] (tensor-core path only): output Hin x Win
};

enum Epi : int {
    EPI_NONE = 0,        // acc + bias
    EPI_RELU = 1,        // relu(acc + bias)
    EPI_GELU = 2,        // exact-erf gelu(acc + bias)
    EPI_RELU_ADDUP = 3,  // relu(acc + bias) + relu(bilinear_x2_align_corners(aux))   (T1 -> R1 input)
    EPI_ADD_RELU = 4,    // relu(acc + bias + aux[m][n])                              (ResNet BasicBlock: + identity; tcgen05 path)
    EPI_ADDUP_RELU = 5,  // relu(acc + bias + bilinear_x2_align_corners(aux))         (FPN fuse conv, low-res half commuted; tcgen05 path)
};

struct GemmDesc {
    const void* A = nullptr;      // activations (float or bf16, per kernel)
    const void* W = nullptr;      // weights (float for SIMT, bf16 for tcgen05)
    const float* bias = nullptr;  // [N] fp32 or null
    void* D = nullptr;            // output (float or bf16)
    const void* aux = nullptr;    // EPI_RELU_ADDUP / EPI_ADDUP_RELU: [img][(Hout/2)*(Wout/2)][N] map; EPI_ADD_RELU: [M][N] (same type as D)
    int M = 0, N = 0, K = 0;
    int64_t lda = 0, ldd = 0;
    int amode = A_PLAIN, epi = EPI_NONE;
    int round_tf32 = 0;           // tcgen05 path, fp32 output: round to the nearest TF32 value (operand of the next tf32 stage)
    // split operands (tcgen05 path): A has only `ka` columns and the k loop re-reads it from column 0 for k >= ka
    // (ka < K <= 2*ka).  Two uses:
    //   K = 2*ka, W = [W_hi | W_lo] (W_lo = rounding residual of W_hi in the operand type): D = A·(W_hi + W_lo)^T, the
    //       weight rounding error is removed at the price of twice the MMAs (free for an HBM-bound stage);
    //   K = 3*k0, ka = 2*k0, A = [x | x_lo], W = [W_hi | W_hi | W_lo]: the "3x" product x_hi·W_hi + x_lo·W_hi + x_hi·W_lo
    //       (x is raw fp32, the tensor core truncates it to TF32 = x_hi; x_lo = tf32(x - x_hi)): fp32-grade token Linears.
    // 0 = plain.
    int ka = 0;
    // pair output (tcgen05 path, fp16 output): the row of D is [hi (N) | lo (N)] with hi = fp16(y), lo = fp16(y - hi): the
    // consumer reads it as the split operand A = [x_hi | x_lo] (ka = 2N) - no activation rounding between the two stages.
    // ldd >= 2N.
    int out_pair = 0;
    // weight-stationary hint (tcgen05 path): keep the group's whole [BN x K] weight slab resident in shared memory and stream
    // A tiles only, when the slab leaves room for >= 4 A stages.  Pays off where the weight re-fetch per tile is a large share of
    // the L2 -> SM traffic: the N = 64 3x3 convs of ResNet layer1 (72 KB slab vs 144 KB of A per tile).
    int ws = 0;
    // A_PLAIN K-split (pose3d flatten "(v c h w)"): k-blocks of kblk elements live kblk_stride apart
    int kblk = 0;
    int64_t kblk_stride = 0;
    // A_CONV3S2 geometry
    int Hin = 0, Win = 0, Cin = 0;   // output is (Hin/2) x (Win/2); M = n_img * Hout * Wout
    // epilogue geometry for EPI_RELU_ADDUP: output positions form Hout x Wout images
    int Hout = 0, Wout = 0;
    // groups
    int groups = 1;
    int64_t a_gs = 0, w_gs = 0, b_gs = 0, d_gs = 0, aux_gs = 0;   // element strides between groups
};

// fp32 SIMT path (reference-grade parity).  TA/TO in {float, bf16}: 0 = float, 1 = bf16.
int gemm_simt(const GemmDesc& d, int a_is_bf16, int d_is_bf16, cudaStream_t st);

// element types of the tcgen05 path's operands / output
enum DT : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
// tcgen05 + TMA path.  A, W both bf16 or both fp16 (kind::f16), or both fp32 read as TF32 (kind::tf32); fp32 accumulate.
// D: fp32, or bf16 / fp16 from 16-bit operands.  fp16 has TF32's 10-bit mantissa at half the bytes and twice the MMA rate; its
// range (|x| < 65504) is the caller's responsibility (conversions saturate).
int gemm_tc(const GemmDesc& d, int in_dt, int out_dt, cudaStream_t st);
int gemm_tc_init();   // resolves cuTensorMapEncodeTiled; EGR_OK or error
// split-K partial sums go to a caller-provided scratch of SPLITK_SCRATCH_BYTES (the engines carve it from their workspace,
// so a forward performs no allocation and can be captured into a CUDA graph); without one (egr_dense_stage) the library
// keeps one lazily allocated scratch per stream.  Thread-local; pass nullptr to clear.
constexpr int64_t SPLITK_SCRATCH_BYTES = 32ll << 20;
void gemm_tc_set_scratch(float* scratch);

struct Up2Coef {
    int i0, i1;
    float l0, l1;
};
// nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True): src = dst * (in-1)/(out-1)
__device__ __forceinline__ Up2Coef up2_coef(int dst, int in_size) {
    const float scale = (float)(in_size - 1) / (float)(2 * in_size - 1);
    const float s = scale * (float)dst;
    Up2Coef c;
    c.i0 = (int)s;
    c.i1 = c.i0 + ((c.i0 < in_size - 1) ? 1 : 0);
    c.l1 = s - (float)c.i0;
    c.l0 = 1.f - c.l1;
    return c;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

}  // namespace egr
