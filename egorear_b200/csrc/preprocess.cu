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
constexpr int PP_THREADS = 1024;

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
// byte `i` (0..3, compile-time) of a word, zero-extended: one PRMT
template <int I> __device__ __forceinline__ int byte_of(uint32_t w) { return (int)__byte_perm(w, 0u, 0x4440u + I); }

struct PpParams {
    const uint8_t* img;           // [N][Hin][Win][3]
    const uint8_t* img_end;       // one past the last byte of the whole input (aligned loads never cross it)
    float* out;                   // [N][3][Hout][Wout]
    uint8_t* out_u8;              // [N][Hout][Wout][3] or null
    const int* bx; const int* kx; const int* by; const int* ky;
    int Hin, Win, Hout, Wout, ksx, ksy, band, rows_max;
    int siw, stw, wpo;            // word strides of the staged / intermediate rows (odd: conflict-free), words per output row
    float mean[3], stdv[3];
};

// aligned 32-bit load that never touches memory at or beyond `end` (only the very last word of the buffer is assembled
// from byte loads)
__device__ __forceinline__ uint32_t load_word_guarded(const uint32_t* p, const uint8_t* end) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
    if (b + 4 <= end) return __ldg(p);
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i)
        if (b + i < end) v |= (uint32_t)b[i] << (8 * i);
    return v;
}

// Shared memory: LUT[3][256] fp32 | staged input rows (row r at word r*siw, byte 0 of the row on a word boundary) |
// 8-bit intermediate rows (word stride stw) | 8-bit output rows of the band.
// Horizontal pass: lane <-> input row, the warp walks the output columns, so bounds and coefficients are warp-uniform
// and the lanes' word loads / byte stores hit 32 different banks (odd row strides).  Vertical pass: lane <-> 4
// consecutive bytes of an intermediate row.  Last pass: one shared-memory LUT lookup per value ((u/255 - mean)/std has
// only 3 x 256 possible results, computed once per CTA with the IEEE ops of the separate torch kernels).
__global__ void __launch_bounds__(PP_THREADS)
preprocess_kernel(const PpParams p) {
    extern __shared__ __align__(16) uint8_t pp_smem[];
    pdl_trigger();
    pdl_wait();
    constexpr int NW = PP_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const int yy0 = blockIdx.x * p.band, yy1 = min(yy0 + p.band, p.Hout);
    const int nb = yy1 - yy0;
    const int r0 = p.by[yy0 * 2];
    const int r1 = p.by[(yy1 - 1) * 2] + p.by[(yy1 - 1) * 2 + 1];      // bounds are monotonic in yy
    const int nrows = r1 - r0;
    const int row_bytes = p.Win * 3;
    float* lut = reinterpret_cast<float*>(pp_smem);
    uint32_t* s_in = reinterpret_cast<uint32_t*>(pp_smem + 3072);
    uint32_t* s_tmp = s_in + (size_t)p.rows_max * p.siw + 4;
    uint32_t* s_out = s_tmp + (size_t)p.rows_max * p.stw;

    for (int i = threadIdx.x; i < 768; i += PP_THREADS) {
        const int c = i >> 8;
        const float mean = (c == 0) ? p.mean[0] : (c == 1) ? p.mean[1] : p.mean[2];
        const float stdv = (c == 0) ? p.stdv[0] : (c == 1) ? p.stdv[1] : p.stdv[2];
        // ToTensor: u / 255 in float32; Normalize: (x - mean) / std, each op rounded like the separate torch kernels
        lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), mean), stdv);
    }
    // ---- stage input rows [r0, r1): warp <-> row, lanes <-> words; re-aligned with a funnel shift when the row does
    //      not start on a word boundary in global memory ----
    const int wpr = (row_bytes + 3) >> 2;
    for (int r = warp; r < nrows; r += NW) {
        const uint8_t* gr = p.img + ((int64_t)n * p.Hin + r0 + r) * row_bytes;
        const int ph = (int)((uintptr_t)gr & 3);
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(gr - ph);
        uint32_t* dst = s_in + (size_t)r * p.siw;
        // words [0, nsafe) of this row lie entirely inside the input buffer (all of them except at the buffer's last word)
        const int nsafe = (int)min((int64_t)wpr + 1, (int64_t)(p.img_end - reinterpret_cast<const uint8_t*>(gw)) >> 2);
        if (ph == 0) {
            const int nfast = min(wpr, nsafe);
#pragma unroll 4
            for (int j = lane; j < nfast; j += 32) dst[j] = __ldg(gw + j);
            for (int j = nfast + lane; j < wpr; j += 32) dst[j] = load_word_guarded(gw + j, p.img_end);
        } else {
            const int nfast = min(wpr, nsafe - 1);               // needs word j + 1 as well
#pragma unroll 4
            for (int j = lane; j < nfast; j += 32) dst[j] = __funnelshift_r(__ldg(gw + j), __ldg(gw + j + 1), 8 * ph);
            for (int j = max(nfast, 0) + lane; j < wpr; j += 32) {
                const uint32_t lo = load_word_guarded(gw + j, p.img_end);
                const uint32_t hi = (4 * j + 4 - ph < row_bytes) ? load_word_guarded(gw + j + 1, p.img_end) : 0u;
                dst[j] = __funnelshift_r(lo, hi, 8 * ph);
            }
        }
    }
    __syncthreads();
    // ---- horizontal pass -> 8-bit intermediate rows (like Pillow's intermediate image) ----
    const int half = 1 << (PP_PRECISION_BITS - 1);
    const int n_rg = (nrows + 31) >> 5;
    uint8_t* s_tmp8 = reinterpret_cast<uint8_t*>(s_tmp);
    for (int it = warp; it < n_rg * p.Wout; it += NW) {
        const int rg = it / p.Wout, xx = it - rg * p.Wout;
        const int r = rg * 32 + lane;
        const int x0 = __ldg(p.bx + xx * 2), cnt = __ldg(p.bx + xx * 2 + 1);
        const int* k = p.kx + xx * p.ksx;
        const int b0 = x0 * 3;
        const int sh = (b0 & 3) * 8;
        const uint32_t* row = s_in + (size_t)min(r, nrows - 1) * p.siw + (b0 >> 2);
        int a0 = half, a1 = half, a2 = half;
        uint32_t w0 = row[0];
        for (int t0 = 0; t0 < cnt; t0 += 4) {            // 4 taps = 12 bytes = 3 re-aligned words per step
            const uint32_t w1 = row[1], w2 = row[2], w3 = row[3];
            row += 3;
            const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
            w0 = w3;
            const int k0 = __ldg(k + t0);
            const int k1 = (t0 + 1 < cnt) ? __ldg(k + t0 + 1) : 0;
            const int k2 = (t0 + 2 < cnt) ? __ldg(k + t0 + 2) : 0;
            const int k3 = (t0 + 3 < cnt) ? __ldg(k + t0 + 3) : 0;
            a0 += byte_of<0>(v0) * k0 + byte_of<3>(v0) * k1 + byte_of<2>(v1) * k2 + byte_of<1>(v2) * k3;
            a1 += byte_of<1>(v0) * k0 + byte_of<0>(v1) * k1 + byte_of<3>(v1) * k2 + byte_of<2>(v2) * k3;
            a2 += byte_of<2>(v0) * k0 + byte_of<1>(v1) * k1 + byte_of<0>(v2) * k2 + byte_of<3>(v2) * k3;
        }
        if (r < nrows) {
            uint8_t* d = s_tmp8 + (size_t)r * p.stw * 4 + xx * 3;
            d[0] = (uint8_t)clip8(a0); d[1] = (uint8_t)clip8(a1); d[2] = (uint8_t)clip8(a2);
        }
    }
    __syncthreads();
    // ---- vertical pass -> 8-bit output rows of the band: warp <-> (row, 32-word chunk), lane <-> 4 bytes ----
    const int nj = (p.wpo + 31) >> 5;
    for (int it = warp; it < nb * nj; it += NW) {
        const int yl = it / nj, j = (it - yl * nj) * 32 + lane;
        const int yy = yy0 + yl;
        const int y0 = __ldg(p.by + yy * 2), cnt = __ldg(p.by + yy * 2 + 1);
        const int* k = p.ky + yy * p.ksy;
        if (j < p.wpo) {
            const uint32_t* src = s_tmp + (size_t)(y0 - r0) * p.stw + j;
            int a0 = half, a1 = half, a2 = half, a3 = half;
            for (int t = 0; t < cnt; ++t) {
                const uint32_t w = src[(size_t)t * p.stw];
                const int c = __ldg(k + t);
                a0 += byte_of<0>(w) * c; a1 += byte_of<1>(w) * c; a2 += byte_of<2>(w) * c; a3 += byte_of<3>(w) * c;
            }
            s_out[yl * p.wpo + j] = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
        }
    }
    __syncthreads();
    // ---- ToTensor + Normalize through the LUT: out[n][c][yy][xx], one coalesced row per warp iteration ----
    const uint8_t* s_out8 = reinterpret_cast<const uint8_t*>(s_out);
    for (int row = warp; row < 3 * nb; row += NW) {
        const int c = row / nb, yl = row - c * nb;
        const uint8_t* src = s_out8 + (size_t)yl * p.wpo * 4 + c;
        float* dst = p.out + (((int64_t)n * 3 + c) * p.Hout + yy0 + yl) * p.Wout;
        const float* l = lut + c * 256;
        for (int xx = lane; xx < p.Wout; xx += 32) __stcs(dst + xx, l[src[xx * 3]]);
    }
    if (p.out_u8) {
        for (int yl = warp; yl < nb; yl += NW) {
            uint8_t* dst = p.out_u8 + ((int64_t)n * p.Hout + yy0 + yl) * p.Wout * 3;
            const uint8_t* src = s_out8 + (size_t)yl * p.wpo * 4;
            for (int i = lane; i < p.Wout * 3; i += 32) dst[i] = src[i];
        }
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
    EGR_CHECK(((uintptr_t)images & 3) == 0, EGR_ERR_INVALID, "preprocess: images must be 4-byte aligned");
    // row strides in words, odd so that 32 lanes on 32 consecutive rows hit 32 different banks
    const int row_bytes = Win * 3;
    const int siw = (((row_bytes + 3) >> 2) + 4) | 1;          // + slack: the tap loop reads up to 4 words past the last tap
    const int wpo = (Wout * 3 + 3) >> 2;
    const int stw = wpo | 1;
    auto rows_for = [&](int band) {
        int mx = 0;
        for (int y0 = 0; y0 < Hout; y0 += band) {
            const int y1 = std::min(y0 + band, Hout) - 1;
            mx = std::max(mx, ty->bounds[y1 * 2] + ty->bounds[y1 * 2 + 1] - ty->bounds[y0 * 2]);
        }
        return mx;
    };
    auto smem_for = [&](int band, int rows) {
        return (size_t)3072 + ((size_t)rows * siw + 4 + (size_t)rows * stw + (size_t)band * wpo) * 4 + 16;
    };
    // band of output rows per CTA: most output rows per 32-row pass of the horizontal stage among those that fit
    int band = 0, rows = 0;
    double best = 0.0;
    for (int b = 1; b <= std::min(Hout, 64); ++b) {
        const int r = rows_for(b);
        if (smem_for(b, r) > (size_t)226 * 1024) continue;
        const double score = (double)b / ((r + 31) / 32);
        if (score > best * 1.0001) { best = score; band = b; rows = r; }
    }
    EGR_CHECK(band > 0, EGR_ERR_UNSUPPORTED, "preprocess: %dx%d -> %dx%d does not fit in shared memory", Hin, Win, Hout, Wout);
    const size_t smem = smem_for(band, rows);
    EGR_CUDA_OK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    PpParams p{};
    p.img = images; p.img_end = images + (size_t)N * Hin * row_bytes; p.out = out; p.out_u8 = resized_u8;
    p.bx = tx->d_bounds; p.kx = tx->d_kk; p.by = ty->d_bounds; p.ky = ty->d_kk;
    p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout; p.ksx = tx->ksize; p.ksy = ty->ksize;
    p.band = band; p.rows_max = rows; p.siw = siw; p.stw = stw; p.wpo = wpo;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean3_host[c]; p.stdv[c] = std3_host[c]; }
    EGR_LAUNCH(preprocess_kernel, dim3((Hout + band - 1) / band, (unsigned)N), PP_THREADS, smem, (cudaStream_t)stream, p);
    return EGR_OK;
}
