// Error plumbing, device check and misc C-ABI entry points.
#include "common.cuh"
#include "gemm.cuh"
#include "layout_ops.cuh"
#include <mutex>
#include <vector>
#include <map>

namespace egr {

std::atomic<int64_t> g_launches{0};

std::string& last_error() {
    static thread_local std::string s;
    return s;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

static int g_dev_state = -1;   // -1 unknown, 0 ok, else error code
static int g_sm_count = 148;
static int g_cc = 0;
static std::mutex g_dev_mu;

int require_device() {
    if (g_dev_state == 0) return EGR_OK;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(EGR_ERR_NO_DEVICE, "egorear_b200 needs a B200 (sm_100) GPU: no CUDA device (%s); there is no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return fail(EGR_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
    g_cc = p.major * 10 + p.minor;
    g_sm_count = p.multiProcessorCount;
    if (p.major != 10) {
        return fail(EGR_ERR_NO_DEVICE, "egorear_b200 is built for sm_100a only; current device '%s' is sm_%d%d", p.name,
                    p.major, p.minor);
    }
    g_dev_state = 0;
    return EGR_OK;
}

int sm_count() { return g_sm_count; }

// ---- stage profiler ----
int g_prof_on = 0;
int g_opt_pdl = 1;
struct Mark { std::string name; cudaEvent_t ev; bool closing; };
static std::vector<Mark> g_marks;
static std::vector<cudaEvent_t> g_ev_pool;
static std::mutex g_prof_mu;

void prof_mark(const char* name, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t ev;
    if (!g_ev_pool.empty()) { ev = g_ev_pool.back(); g_ev_pool.pop_back(); }
    else if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    g_marks.push_back(Mark{name ? name : "", ev, name == nullptr});
}

}  // namespace egr

extern "C" const char* egr_last_error(void) { return egr::last_error().c_str(); }
extern "C" int egr_version(void) { return 100; }
extern "C" int64_t egr_launch_count(void) { return egr::g_launches.load(); }
extern "C" int egr_device_check(int* cc, int* sms) {
    int rc = egr::require_device();
    if (cc) *cc = egr::g_cc;
    if (sms) *sms = egr::g_sm_count;
    return rc;
}

extern "C" int egr_profile_enable(int on) {
    egr::g_prof_on = on ? 1 : 0;
    return EGR_OK;
}

// Aggregates the recorded intervals as "name:total_ms:count;" pairs into buf (NUL terminated), clears the marks.
// Synchronises on the last recorded event.
extern "C" int egr_profile_read(char* buf, int cap) {
    using namespace egr;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<double, int>> agg;
    std::vector<std::string> order;
    if (!g_marks.empty()) cudaEventSynchronize(g_marks.back().ev);
    for (size_t i = 0; i + 1 < g_marks.size(); ++i) {
        if (g_marks[i].closing) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_marks[i].ev, g_marks[i + 1].ev) != cudaSuccess) { cudaGetLastError(); continue; }
        auto it = agg.find(g_marks[i].name);
        if (it == agg.end()) { agg[g_marks[i].name] = std::make_pair((double)ms, 1); order.push_back(g_marks[i].name); }
        else { it->second.first += ms; it->second.second += 1; }
    }
    for (auto& m : g_marks) g_ev_pool.push_back(m.ev);
    g_marks.clear();
    std::string out;
    char tmp[256];
    for (auto& n : order) {
        snprintf(tmp, sizeof(tmp), "%s:%.6f:%d;", n.c_str(), agg[n].first, agg[n].second);
        out += tmp;
    }
    if (!buf || cap <= 0) return EGR_OK;
    if ((int)out.size() >= cap) return egr::fail(EGR_ERR_INVALID, "profile_read: buffer too small (%d needed)", (int)out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return EGR_OK;
}

extern "C" int egr_dense_stage(const egr_dense_desc* c, void* stream) {
    using namespace egr;
    EGR_CHECK(c && c->A && c->W && c->D, EGR_ERR_INVALID, "dense_stage: null descriptor / operand");
    if (int rc = require_device()) return rc;
    gemm_tc_set_scratch(nullptr);       // standalone stage: library-owned split-K scratch of this stream
    GemmDesc d;
    d.A = c->A; d.W = c->W; d.bias = c->bias; d.D = c->D; d.aux = c->aux;
    d.M = c->M; d.N = c->N; d.K = c->K; d.lda = c->lda; d.ldd = c->ldd; d.amode = c->amode; d.epi = c->epi;
    d.kblk = c->kblk; d.kblk_stride = c->kblk_stride; d.ka = c->ka; d.out_pair = c->out_pair;
    EGR_CHECK(d.ka == 0 || c->use_tc, EGR_ERR_UNSUPPORTED, "dense_stage: split operands (ka) need the tensor-core kernel");
    d.Hin = c->Hin; d.Win = c->Win; d.Cin = c->Cin; d.Hout = c->Hout; d.Wout = c->Wout;
    d.groups = c->groups > 0 ? c->groups : 1;
    d.a_gs = c->a_gs; d.w_gs = c->w_gs; d.b_gs = c->b_gs; d.d_gs = c->d_gs; d.aux_gs = c->aux_gs;
    EGR_CHECK(d.amode >= A_PLAIN && d.amode <= A_CONV3S1, EGR_ERR_INVALID, "dense_stage: amode %d", d.amode);
    EGR_CHECK(d.epi >= EPI_NONE && d.epi <= EPI_ADDUP_RELU, EGR_ERR_INVALID, "dense_stage: epi %d", d.epi);
    cudaStream_t st = (cudaStream_t)stream;
    if (c->use_tc) {
        // a_is_bf16 / d_is_bf16: 0 fp32, 1 bf16, 2 fp16
        return gemm_tc(d, c->a_is_bf16, c->d_is_bf16, st);
    }
    return gemm_simt(d, c->a_is_bf16, c->d_is_bf16, st);
}

extern "C" int egr_head_tail_stage(const void* z, const float* w, const float* bias, const int32_t* wsel, int B, int G, int J,
                                   float* hm, int64_t hm_bs, int64_t hm_gs, void* hm_t, int precise, int impl, void* stream) {
    using namespace egr;
    EGR_CHECK(z && w && bias && wsel && hm, EGR_ERR_INVALID, "head_tail_stage: null operand");
    EGR_CHECK(B > 0 && G > 0 && G <= 4 && J > 0 && J <= 16, EGR_ERR_INVALID, "head_tail_stage: B=%d G=%d J=%d", B, G, J);
    EGR_CHECK((uintptr_t)z % 16 == 0 && (uintptr_t)w % 16 == 0, EGR_ERR_INVALID, "head_tail_stage: z and w must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    int sel[4] = {0, 0, 0, 0};
    for (int g = 0; g < G; ++g) sel[g] = wsel[g];
    cudaStream_t st = (cudaStream_t)stream;
    if (impl) return head_tail_mma(z, w, bias, sel, B, G, J, hm, hm_bs, hm_gs, hm_t, precise, st);
    return head_tail_tc(z, w, bias, sel, B, G, J, hm, hm_bs, hm_gs, hm_t, precise, st);
}

extern "C" int egr_up2_relu_stage(const void* z, int n_img, void* out, int out_f16, int impl, void* stream) {
    using namespace egr;
    EGR_CHECK(z && out && n_img > 0, EGR_ERR_INVALID, "up2_relu_stage: null operand / n_img=%d", n_img);
    EGR_CHECK((uintptr_t)z % 16 == 0 && (uintptr_t)out % 16 == 0, EGR_ERR_INVALID, "up2_relu_stage: z and out must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl) return up2_relu_mma(z, n_img, out, out_f16, st);
    return up2_relu_dual(z, 2, n_img, 1, 32, 32, 128, nullptr, 0, 0, out, out_f16 ? 2 : 1, nullptr, 0, st, out_f16 ? 1 : 0,
                         /*allow_mma=*/false);
}
