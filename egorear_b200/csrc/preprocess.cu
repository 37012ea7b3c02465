// Image preprocessing of the reference's datasets on the GPU:
//   Image.resize([256, 256], Image.BICUBIC) -> ToTensor -> Normalize(mean, std)
//   (pose_estimation/datasets/ego4view_rw/ego4view_rw_heatmap_mvf.py:40-41,96-99 and the five sibling dataset classes)
// bit-exact with Pillow + torchvision.  Pillow's resampler (src/libImaging/Resample.c; Pillow is an un-vendored
// dependency, README.md:133) works on 8-bit data in two separable passes with an 8-bit intermediate image and 22-bit
// fixed-point coefficients: the coefficient tables are computed on the host in double exactly as precompute_coeffs /
// normalize_coeffs_8bpc do, the kernel does the integer MACs, the clip to 8 bits, and the float32 (x/255 - mean)/std.
//
// One CTA = one band of output rows of one image: the input rows the band needs are staged in shared memory with
// 128-bit loads (they are one contiguous byte range), the horizontal pass writes the 8-bit intermediate rows to shared
// memory, the vertical pass reads them and writes the normalised NCHW fp32 rows coalesced.  HBM traffic = the decoded
// image once (neighbouring bands re-read their halo rows from L2) + the output once.
#include "common.cuh"
#include <mutex>
#include <map>
#include <vector>
#include <math.h>
#include <algorithm>

namespace egr {
namespace {

constexpr int PP_PRECISION_BITS = 32 - 8 - 2;
constexpr int PP_THREADS = 512;

double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

struct ResampleTable {
    int ksize = 0;
    std::vector<int> bounds;      // [out][2] = (first input index, tap count)
    std::vector<int> kk;          // [out][ksize], fixed point
    int* d_bounds = nullptr;
    int* d_kk = nullptr;
};

// Resample.c precompute_coeffs (box = the whole image) + normalize_coeffs_8bpc
void build_table(int in_size, int out_size, ResampleTable& t) {
    double scale, filterscale;
    filterscale = scale = (double)((float)in_size - 0.0f) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    t.ksize = ksize;
    t.bounds.assign((size_t)out_size * 2, 0);
    t.kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < out_size; xx++) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; x++) {
            const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; x++) {
            if (ww != 0.0) k[x] /= ww;
            t.kk[(size_t)xx * ksize + x] = (k[x] < 0) ? (int)(-0.5 + k[x] * (1 << PP_PRECISION_BITS))
                                                      : (int)(0.5 + k[x] * (1 << PP_PRECISION_BITS));
        }
        t.bounds[xx * 2 + 0] = xmin;
        t.bounds[xx * 2 + 1] = xmax;
    }
}

std::mutex g_tab_mu;
std::map<std::tuple<int, int, int>, ResampleTable> g_tabs;     // (device, in, out)

int get_table(int in_size, int out_size, const ResampleTable** out) {
    int dev = 0;
    EGR_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_tuple(dev, in_size, out_size);
    auto it = g_tabs.find(key);
    if (it == g_tabs.end()) {
        ResampleTable t;
        build_table(in_size, out_size, t);
        EGR_CUDA_OK(cudaMalloc(&t.d_bounds, t.bounds.size() * sizeof(int)));
        EGR_CUDA_OK(cudaMalloc(&t.d_kk, t.kk.size() * sizeof(int)));
        EGR_CUDA_OK(cudaMemcpy(t.d_bounds, t.bounds.data(), t.bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
        EGR_CUDA_OK(cudaMemcpy(t.d_kk, t.kk.data(), t.kk.size() * sizeof(int), cudaMemcpyHostToDevice));
        it = g_tabs.emplace(key, std::move(t)).first;
    }
    *out = &it->second;
    return EGR_OK;
}

__device__ __forceinline__ int clip8(int v) { return min(max(v >> PP_PRECISION_BITS, 0), 255); }

struct PpParams {
    const uint8_t* img;           // [N][Hin][Win][3]
    float* out;                   // [N][3][Hout][Wout]
    uint8_t* out_u8;              // [N][Hout][Wout][3] or null
    const int* bx; const int* kx; const int* by; const int* ky;
    int Hin, Win, Hout, Wout, ksx, ksy, band, rows_max;
    float mean[3], stdv[3];
};

__global__ void __launch_bounds__(PP_THREADS)
preprocess_kernel(const PpParams p) {
    extern __shared__ __align__(16) uint8_t pp_smem[];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y;
    const int yy0 = blockIdx.x * p.band, yy1 = min(yy0 + p.band, p.Hout);
    const int r0 = p.by[yy0 * 2];
    const int r1 = p.by[(yy1 - 1) * 2] + p.by[(yy1 - 1) * 2 + 1];      // bounds are monotonic in yy
    const int nrows = r1 - r0;
    const int row_bytes = p.Win * 3;
    // ---- stage the contiguous byte range of input rows [r0, r1) with the same 16-byte phase as in global memory ----
    const uint8_t* g0 = p.img + ((int64_t)n * p.Hin + r0) * row_bytes;
    const int phase = (int)((uintptr_t)g0 & 15);
    uint8_t* s_in = pp_smem + phase;
    uint8_t* s_tmp = pp_smem + (((size_t)p.rows_max * row_bytes + 32 + 15) & ~(size_t)15);
    const int total = nrows * row_bytes;
    const int head = min((16 - phase) & 15, total);
    const int n16 = (total - head) >> 4;
    const int tail0 = head + (n16 << 4);
    if ((int)threadIdx.x < head) s_in[threadIdx.x] = g0[threadIdx.x];
    {
        const uint4* gs = reinterpret_cast<const uint4*>(g0 + head);
        uint4* ss = reinterpret_cast<uint4*>(s_in + head);
        int i = threadIdx.x;
        for (; i + 3 * PP_THREADS < n16; i += 4 * PP_THREADS) {          // 4 independent 128-bit loads per thread
            const uint4 a = __ldg(gs + i), b = __ldg(gs + i + PP_THREADS), c = __ldg(gs + i + 2 * PP_THREADS),
                        d = __ldg(gs + i + 3 * PP_THREADS);
            ss[i] = a; ss[i + PP_THREADS] = b; ss[i + 2 * PP_THREADS] = c; ss[i + 3 * PP_THREADS] = d;
        }
        for (; i < n16; i += PP_THREADS) ss[i] = __ldg(gs + i);
    }
    if ((int)threadIdx.x < total - tail0) s_in[tail0 + threadIdx.x] = g0[tail0 + threadIdx.x];
    __syncthreads();
    // ---- horizontal pass: s_in rows -> s_tmp [nrows][Wout][3] (8-bit, like Pillow's intermediate image) ----
    const int half = 1 << (PP_PRECISION_BITS - 1);
    for (int idx = threadIdx.x; idx < nrows * p.Wout; idx += PP_THREADS) {
        const int r = idx / p.Wout, xx = idx - r * p.Wout;
        const int x0 = __ldg(p.bx + xx * 2), cnt = __ldg(p.bx + xx * 2 + 1);
        const int* k = p.kx + xx * p.ksx;
        const uint8_t* src = s_in + (r * p.Win + x0) * 3;
        int a0 = half, a1 = half, a2 = half;
        for (int t = 0; t < cnt; ++t) {
            const int c = __ldg(k + t);
            a0 += (int)src[t * 3 + 0] * c;
            a1 += (int)src[t * 3 + 1] * c;
            a2 += (int)src[t * 3 + 2] * c;
        }
        uint8_t* d = s_tmp + idx * 3;
        d[0] = (uint8_t)clip8(a0); d[1] = (uint8_t)clip8(a1); d[2] = (uint8_t)clip8(a2);
    }
    __syncthreads();
    // ---- vertical pass + ToTensor + Normalize: out[n][c][yy][xx], xx fastest ----
    const int nb = yy1 - yy0;
    for (int idx = threadIdx.x; idx < 3 * nb * p.Wout; idx += PP_THREADS) {
        const int c = idx / (nb * p.Wout);
        const int rem = idx - c * nb * p.Wout;
        const int yl = rem / p.Wout, xx = rem - yl * p.Wout;
        const int yy = yy0 + yl;
        const int y0 = __ldg(p.by + yy * 2), cnt = __ldg(p.by + yy * 2 + 1);
        const int* k = p.ky + yy * p.ksy;
        const uint8_t* src = s_tmp + ((y0 - r0) * p.Wout + xx) * 3 + c;
        int a = half;
        for (int t = 0; t < cnt; ++t) a += (int)src[t * p.Wout * 3] * __ldg(k + t);
        const int u = clip8(a);
        if (p.out_u8) p.out_u8[(((int64_t)n * p.Hout + yy) * p.Wout + xx) * 3 + c] = (uint8_t)u;
        const float mean = (c == 0) ? p.mean[0] : (c == 1) ? p.mean[1] : p.mean[2];
        const float stdv = (c == 0) ? p.stdv[0] : (c == 1) ? p.stdv[1] : p.stdv[2];
        // ToTensor: u / 255 in float32; Normalize: (x - mean) / std, each op rounded like the separate torch kernels
        const float x = __fdiv_rn((float)u, 255.0f);
        __stcs(p.out + (((int64_t)n * 3 + c) * p.Hout + yy) * p.Wout + xx, __fdiv_rn(__fsub_rn(x, mean), stdv));
    }
}

}  // namespace
}  // namespace egr

using namespace egr;

// host-only: the coefficient table the kernel uses (no device needed; tests compare it with the oracle's)
extern "C" int egr_resample_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk) {
    EGR_CHECK(in_size > 0 && out_size > 0 && ksize, EGR_ERR_INVALID, "resample_coeffs: sizes must be positive");
    ResampleTable t;
    build_table(in_size, out_size, t);
    *ksize = t.ksize;
    if (bounds) std::copy(t.bounds.begin(), t.bounds.end(), bounds);
    if (kk) std::copy(t.kk.begin(), t.kk.end(), kk);
    return EGR_OK;
}

extern "C" int egr_preprocess_images(const uint8_t* images, int64_t N, int Hin, int Win, int Hout, int Wout,
                                     const float* mean3_host, const float* std3_host, float* out, uint8_t* resized_u8,
                                     void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(N >= 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, EGR_ERR_INVALID, "preprocess: images should be [N, H, W, 3] uint8");
    EGR_CHECK(N <= 65535, EGR_ERR_UNSUPPORTED, "preprocess: at most 65535 images per call");
    EGR_CHECK(mean3_host && std3_host, EGR_ERR_INVALID, "preprocess: mean / std are required");
    if (N == 0) return EGR_OK;
    EGR_CHECK(images && out, EGR_ERR_INVALID, "preprocess: null pointer");
    const ResampleTable *tx = nullptr, *ty = nullptr;
    if (int rc = get_table(Win, Wout, &tx)) return rc;
    if (int rc = get_table(Hin, Hout, &ty)) return rc;
    // largest band of output rows whose input rows + intermediate rows fit in shared memory
    const int row_bytes = Win * 3;
    auto rows_for = [&](int band) {
        int mx = 0;
        for (int y0 = 0; y0 < Hout; y0 += band) {
            const int y1 = std::min(y0 + band, Hout) - 1;
            mx = std::max(mx, ty->bounds[y1 * 2] + ty->bounds[y1 * 2 + 1] - ty->bounds[y0 * 2]);
        }
        return mx;
    };
    auto smem_for = [&](int rows) { return (((size_t)rows * row_bytes + 32 + 15) & ~(size_t)15) + (size_t)rows * Wout * 3 + 16; };
    int band = 0, rows = 0;
    for (int b : {16, 8, 4, 2, 1}) {          // fewest re-staged halo rows: the largest band that fits
        const int r = rows_for(b);
        if (smem_for(r) <= (size_t)200 * 1024) { band = b; rows = r; break; }
    }
    EGR_CHECK(band > 0, EGR_ERR_UNSUPPORTED, "preprocess: %dx%d -> %dx%d does not fit in shared memory", Hin, Win, Hout, Wout);
    const size_t smem = smem_for(rows);
    EGR_CUDA_OK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PpParams p{};
    p.img = images; p.out = out; p.out_u8 = resized_u8;
    p.bx = tx->d_bounds; p.kx = tx->d_kk; p.by = ty->d_bounds; p.ky = ty->d_kk;
    p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout; p.ksx = tx->ksize; p.ksy = ty->ksize;
    p.band = band; p.rows_max = rows;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean3_host[c]; p.stdv[c] = std3_host[c]; }
    EGR_LAUNCH(preprocess_kernel, dim3((Hout + band - 1) / band, (unsigned)N), PP_THREADS, smem, (cudaStream_t)stream, p);
    return EGR_OK;
}
