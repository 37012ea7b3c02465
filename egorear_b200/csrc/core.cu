// Error plumbing, device check and misc C-ABI entry points.
#include "common.cuh"
#include <mutex>

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
