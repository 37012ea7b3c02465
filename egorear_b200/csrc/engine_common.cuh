// Pieces shared by the two engines: parameter registry, library-owned device buffers, workspace carving.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "layout_ops.cuh"
#include <string>
#include <unordered_map>
#include <vector>

namespace egr {

struct ParamRef {
    const float* ptr = nullptr;
    int64_t numel = 0;
};

struct ParamTable {
    std::unordered_map<std::string, ParamRef> map;
    int set(const char* key, const float* ptr, int64_t numel) {
        if (!key || !ptr || numel <= 0) return fail(EGR_ERR_INVALID, "set_param: bad argument for '%s'", key ? key : "(null)");
        map[key] = ParamRef{ptr, numel};
        return EGR_OK;
    }
    // fetch with a size check; returns nullptr and sets the error on failure
    const float* get(const std::string& key, int64_t numel, int* rc) const {
        auto it = map.find(key);
        if (it == map.end()) {
            *rc = fail(EGR_ERR_STATE, "missing parameter '%s' (state_dict key not registered)", key.c_str());
            return nullptr;
        }
        if (it->second.numel != numel) {
            *rc = fail(EGR_ERR_INVALID, "parameter '%s' has %lld elements, expected %lld", key.c_str(),
                       (long long)it->second.numel, (long long)numel);
            return nullptr;
        }
        return it->second.ptr;
    }
};

// library-owned device allocations of a handle (derived weights); freed with the handle / on re-prepack
struct DevPool {
    std::vector<void*> blocks;
    int64_t bytes = 0;
    template <typename T>
    int alloc(T** out, int64_t n) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, (size_t)(n * (int64_t)sizeof(T)));
        if (e != cudaSuccess) return fail(EGR_ERR_CUDA, "cudaMalloc(%lld B) failed: %s", (long long)(n * sizeof(T)), cudaGetErrorString(e));
        blocks.push_back(p);
        bytes += n * (int64_t)sizeof(T);
        *out = reinterpret_cast<T*>(p);
        return EGR_OK;
    }
    void release() {
        for (void* p : blocks) cudaFree(p);
        blocks.clear();
        bytes = 0;
    }
};

// bump allocator over the caller-provided workspace
struct Carver {
    char* base;
    int64_t off = 0, cap;
    Carver(void* b, int64_t c) : base(reinterpret_cast<char*>(b)), cap(c) {}
    void* take(int64_t bytes) {
        off = (off + 255) & ~int64_t(255);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
    bool ok() const { return off <= cap; }
};

// a weight matrix [N][K] kept as fp32 and (bf16 mode) as a bf16 copy, `sets` weight sets at constant stride
struct WMat {
    float* f32 = nullptr;
    __nv_bfloat16* bf16 = nullptr;
    __half* f16 = nullptr;
    float* bias = nullptr;    // [sets][N]
    int N = 0, K = 0, sets = 0;
    int kw = 1;               // EGR_PREC_FP16: f16 holds [N][kw*K] = [W_hi | W_lo] when kw == 2 (split weights, gemm.cuh `ka`)
    // fp32-grade token Linears (EGR_PREC_FP16): f32x3 [sets][N][3K] = [W_hi | W_hi | W_lo] in TF32 values; A is [x | x_lo]
    float* f32x3 = nullptr;
    __half* f16x3 = nullptr;  // the same product on fp16 pairs: [sets][N][3K] = [W_hi | W_hi | W_lo] fp16, A rows [x_hi | x_lo] fp16
    const void* w(int prec_bf16_tc) const { return prec_bf16_tc ? (const void*)bf16 : (const void*)f32; }
    int64_t stride() const { return (int64_t)N * K; }
};

extern int g_opt_tc;   // 1: bf16 precision uses the tcgen05 kernel; 0: bf16 activations through the SIMT kernel (debug)
extern int g_opt_wsplit;   // EGR_PREC_FP16: 0 single fp16 weights, 1 (default) hi + lo pairs for 1x1 / Linear weights, 2 also the 3x3 convs
extern int g_opt_tok3x;    // EGR_PREC_FP16: mvfex token Linears as "3x TF32" (default 1)

inline bool is16(int prec) { return prec == EGR_PREC_BF16 || prec == EGR_PREC_FP16; }
inline int act_code(int prec) { return prec == EGR_PREC_BF16 ? 1 : prec == EGR_PREC_FP16 ? 2 : 0; }     // 0 fp32, 1 bf16, 2 fp16
inline int stage_mode(int prec) { return prec == EGR_PREC_FP16 ? 3 : prec == EGR_PREC_BF16 ? 1 : 0; }    // nchw_to_nhwc out_mode

// internal stage precision on top of the public EGR_PREC_*: fp32 activations and weights through the tensor cores as
// TF32 (10-bit mantissa, fp32 accumulate) — used where bf16 operands would eat the 0.1 mm MPJPE budget (pose3d P2)
constexpr int PREC_TF32 = 2;
// fp16 activations and weights (kind::f16): TF32's 10-bit mantissa at half the bytes and twice the MMA rate; range
// |x| < 65504 (conversions saturate) — the default of the pose3d proposal branch in EGR_PREC_BF16 (option "pose_p2_fp16")
constexpr int PREC_FP16 = 3;      // == EGR_PREC_FP16 (public since round 2: every 16-bit dense stage in fp16, split 1x1 weights)
// token Linears of EGR_PREC_FP16: "3x TF32" (see gemm.cuh `ka`), A buffers hold [x | x_lo] rows of 2K floats
constexpr int PREC_TF32X3 = 4;
// the folded memory-projection GEMM (K = 3200, by far the largest token GEMM) runs the same three-term product on fp16 pairs:
// 21 mantissa bits per operand at half the L2 -> smem traffic and twice the MMA rate; fp32 output
constexpr int PREC_F16X3 = 5;

// run one dense stage in the handle's precision.  `out_f32`: the output stays fp32 even in bf16 mode.
// `out16`: DT_F16 makes a bf16-mode tensor-core stage write fp16 instead of bf16 (interpolated right afterwards in half2).
inline int run_gemm(GemmDesc d, const WMat& w, int set_begin, int prec, bool out_f32, cudaStream_t st, int out16 = DT_BF16) {
    const bool bf = (prec == EGR_PREC_BF16);
    const bool tc = bf && g_opt_tc;
    d.N = w.N;
    d.K = w.K;
    if (prec == PREC_FP16) {
        d.K = w.K * w.kw;
        d.ka = (w.kw > 1) ? w.K : 0;
        d.W = w.f16 + (int64_t)set_begin * w.stride() * w.kw;
        d.bias = w.bias ? w.bias + (int64_t)set_begin * w.N : nullptr;
        d.w_gs = w.stride() * w.kw;
        d.b_gs = w.N;
        return gemm_tc(d, DT_F16, out_f32 ? DT_F32 : DT_F16, st);
    }
    if (prec == PREC_TF32X3) {      // caller: d.lda = 2K (rows [x | x_lo])
        d.K = 3 * w.K;
        d.ka = 2 * w.K;
        d.W = w.f32x3 + (int64_t)set_begin * w.stride() * 3;
        d.bias = w.bias ? w.bias + (int64_t)set_begin * w.N : nullptr;
        d.w_gs = w.stride() * 3;
        d.b_gs = w.N;
        return gemm_tc(d, DT_F32, DT_F32, st);
    }
    if (prec == PREC_F16X3) {       // caller: d.A fp16 rows [x_hi | x_lo] (a pair output of the previous stage), d.lda = 2K
        d.K = 3 * w.K;
        d.ka = 2 * w.K;
        d.W = w.f16x3 + (int64_t)set_begin * w.stride() * 3;
        d.bias = w.bias ? w.bias + (int64_t)set_begin * w.N : nullptr;
        d.w_gs = w.stride() * 3;
        d.b_gs = w.N;
        return gemm_tc(d, DT_F16, out_f32 ? DT_F32 : DT_F16, st);
    }
    if (prec == PREC_TF32 && g_opt_tc) {
        d.W = w.f32 + (int64_t)set_begin * w.stride();
        d.bias = w.bias ? w.bias + (int64_t)set_begin * w.N : nullptr;
        d.w_gs = w.stride();
        d.b_gs = w.N;
        return gemm_tc(d, DT_F32, DT_F32, st);
    }
    d.W = tc ? (const void*)(w.bf16 + (int64_t)set_begin * w.stride())
             : (const void*)(w.f32 + (int64_t)set_begin * w.stride());
    d.bias = w.bias ? w.bias + (int64_t)set_begin * w.N : nullptr;
    d.w_gs = w.stride();
    d.b_gs = w.N;
    const int d_bf16 = (bf && !out_f32) ? 1 : 0;
    if (tc) return gemm_tc(d, DT_BF16, d_bf16 ? out16 : (int)DT_F32, st);
    return gemm_simt(d, bf ? 1 : 0, d_bf16, st);
}

}  // namespace egr
