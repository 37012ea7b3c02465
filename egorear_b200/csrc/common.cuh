// Shared helpers for the egorear_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <atomic>

#include "../../include/egorear_b200.h"

namespace egr {

// ---- error plumbing -------------------------------------------------------------------------
std::string& last_error();                       // thread-local
int fail(int code, const char* fmt, ...);        // formats into last_error(), returns code
extern std::atomic<int64_t> g_launches;          // kernels launched by this library

#define EGR_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::egr::fail(EGR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                \
    } while (0)

#define EGR_CHECK(cond, code, ...)                                                                 \
    do {                                                                                           \
        if (!(cond)) return ::egr::fail((code), __VA_ARGS__);                                      \
    } while (0)

// count + check a kernel launch
#define EGR_LAUNCHED()                                                                             \
    do {                                                                                           \
        ::egr::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return ::egr::fail(EGR_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                \
    } while (0)

// Programmatic dependent launch (PDL): every hot-path kernel starts with pdl_trigger() (the next kernel of the stream may
// begin launching: its CTAs become resident and run their prologue as this grid drains) and executes pdl_wait() before its
// first global-memory access (returns once ALL earlier grids have completed and flushed).  Because every kernel waits
// before touching memory, ordering is exactly that of plain stream serialisation; only launch latency and prologues
// (barrier init, TMEM allocation, descriptor prefetch) overlap.  Kernels that allocate TMEM trigger only AFTER their own
// allocation, so a dependent CTA can never hold TMEM columns that an unfinished primary CTA still needs.
extern int g_opt_pdl;
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_opt_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// launch + count + check; a templated kernel name goes in parentheses: EGR_LAUNCH((k<A, B>), grid, block, smem, st, args...)
#define EGR_LAUNCH(kernel, grid, block, smem, st, ...)                                             \
    do {                                                                                           \
        cudaError_t _e = ::egr::launch_k(kernel, dim3(grid), dim3(block), (size_t)(smem), (st), __VA_ARGS__); \
        ::egr::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
        if (_e == cudaSuccess) _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess)                                                                     \
            return ::egr::fail(EGR_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                \
    } while (0)

// Stage profiler (bench.py roofline): when enabled, prof_mark(name, stream) records a CUDA event on the launch
// stream; the time between a mark and the next one is attributed to `name` (nullptr closes the last interval).
void prof_mark(const char* name, cudaStream_t st);
extern int g_prof_on;
#define EGR_MARK(name, st) do { if (::egr::g_prof_on) ::egr::prof_mark((name), (st)); } while (0)

int require_device();   // EGR_OK iff current device is sm_100; caches the answer
int sm_count();

// ---- device helpers -------------------------------------------------------------------------
template <typename T> struct ActT;
template <> struct ActT<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct ActT<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct ActT<__half> {
    static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void st(__half* p, float v) {       // saturating (|x| > 65504 -> +-65504)
        unsigned short h;
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
        *reinterpret_cast<unsigned short*>(p) = h;
    }
};

// fp32 -> nearest TF32 value (10-bit mantissa, ties away), kept in an fp32 container.  The tensor core TRUNCATES fp32
// operands to TF32; pre-rounding makes the operand error unbiased (measured: truncation alone costs 0.12 mm MPJPE).
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace egr
