// TEST INFRASTRUCTURE: compiles the per-sample arithmetic of egorear_b200/csrc/eval_pose_math.cuh for the HOST (g++),
// so that tests/test_eval_metrics_host.py can check the 3x3 SVD / Procrustes / PCK logic against numpy on a machine
// without a GPU.  Not part of libegorear_b200.so; the product path runs the same header inside eval_metrics.cu.
#define EGR_HD inline
#include "../../egorear_b200/csrc/eval_pose_math.cuh"
#include <stdint.h>

namespace {
struct Rows {
    const float* p;
    float operator()(int j, int c) const { return p[j * 3 + c]; }
};
}  // namespace

extern "C" void eval_pose_host(const float* pred, const float* gt, int64_t B, int J, float unit_scale, float pck_thr,
                               const float* auc_thr, int n_auc, double* metrics, double* s1_hat) {
    bool sorted = true;
    for (int i = 0; i + 1 < n_auc; ++i) sorted = sorted && (auc_thr[i] <= auc_thr[i + 1]);
    for (int64_t b = 0; b < B; ++b) {
        Rows p{pred + b * J * 3}, g{gt + b * J * 3};
        egr::eval_pose_sample(p, g, J, unit_scale, pck_thr, auc_thr, n_auc, sorted, metrics + b * 4,
                              s1_hat ? s1_hat + b * J * 3 : nullptr);
    }
}
