// Per-sample arithmetic of the eval-time pose metrics, shared by the CUDA kernel
// (eval_metrics.cu) and by a host-compiled test harness (tests/host/eval_pose_host.cpp, EGR_HD empty) so that the
// 3x3 SVD / Procrustes logic is checked against numpy on machines without a GPU.
//
// Reference: pose_estimation/models/utils/pose_metric.py:120-167 (compute_similarity_transform, per-sample numpy;
// float32 inputs, float64 from `Z = np.eye(3)` on), pose_estimation/utils/loss.py:9-48 (MPJPE / PCK / AUC).
#pragma once
#include <math.h>

#ifndef EGR_HD
#define EGR_HD __host__ __device__ __forceinline__
#endif

namespace egr {

struct SimTransform {
    double sR[9];     // scale * R, row-major: S1_hat = sR * x + t
    double t[3];
};

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix held column-wise: on exit the columns of `a` are U*diag(sigma)
// and `v` holds V (K = U diag(sigma) V^T).  Quadratically convergent; 3x3 needs <= 6 sweeps in double.
EGR_HD void jacobi_svd3(double a[3][3] /*[col][row]*/, double v[3][3] /*[col][row]*/) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) v[c][r] = (c == r) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pair = 0; pair < 3; ++pair) {
            const int p = (pair == 2) ? 1 : 0;
            const int q = (pair == 0) ? 1 : 2;
            const double alpha = a[p][0] * a[p][0] + a[p][1] * a[p][1] + a[p][2] * a[p][2];
            const double beta = a[q][0] * a[q][0] + a[q][1] * a[q][1] + a[q][2] * a[q][2];
            const double gamma = a[p][0] * a[q][0] + a[p][1] * a[q][1] + a[p][2] * a[q][2];
            const double lim = 1e-15 * sqrt(alpha * beta);
            if (fabs(gamma) > lim && gamma != 0.0) {
                off += fabs(gamma);
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt);
                const double sn = cs * tt;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double ap = a[p][r], aq = a[q][r];
                    a[p][r] = cs * ap - sn * aq;
                    a[q][r] = sn * ap + cs * aq;
                    const double vp = v[p][r], vq = v[q][r];
                    v[p][r] = cs * vp - sn * vq;
                    v[q][r] = sn * vp + cs * vq;
                }
            }
        }
        if (off == 0.0) break;
    }
}

EGR_HD double det3(const double m[3][3]) {
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

// Similarity transform (s, R, t) taking the J points `s1` onto `s2` (both [J,3] fp32, row stride 3):
// mu, X = S - mu, var1 = sum X1^2, K = X1^T-outer-X2, K = U S V^T, Z = diag(1,1,sign det(U V^T)) on the SMALLEST
// singular value (LAPACK orders them descending, the reference flips the last one), R = V Z U^T,
// s = trace(R K) / var1 = sum z_i sigma_i / var1, t = mu2 - s R mu1.
template <class Load>
EGR_HD void similarity_transform(Load s1, Load s2, int J, SimTransform& out) {
    double mu1[3] = {0, 0, 0}, mu2[3] = {0, 0, 0};
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) { mu1[c] += (double)s1(j, c); mu2[c] += (double)s2(j, c); }
#pragma unroll
    for (int c = 0; c < 3; ++c) { mu1[c] /= (double)J; mu2[c] /= (double)J; }
    double var1 = 0.0;
    double k[3][3];                  // k[col][row] = K[row][col], K = X1 (3xJ) * X2^T (Jx3)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) k[c][r] = 0.0;
    for (int j = 0; j < J; ++j) {
        double x1[3], x2[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { x1[c] = (double)s1(j, c) - mu1[c]; x2[c] = (double)s2(j, c) - mu2[c]; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            var1 += x1[c] * x1[c];
#pragma unroll
            for (int r = 0; r < 3; ++r) k[c][r] += x1[r] * x2[c];
        }
    }
    double v[3][3];
    jacobi_svd3(k, v);
    double sig[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) sig[c] = sqrt(k[c][0] * k[c][0] + k[c][1] * k[c][1] + k[c][2] * k[c][2]);
    // order (descending) without moving data: i0 largest, i2 smallest
    int i0 = 0, i2 = 2;
    if (sig[1] > sig[i0]) i0 = 1;
    if (sig[2] > sig[i0]) i0 = 2;
    i2 = (i0 == 2) ? 1 : 2;
    { const int o = 3 - i0 - i2; if (sig[o] < sig[i2]) i2 = o; }
    const int i1 = 3 - i0 - i2;
    double u[3][3];                  // u[col][row]
    const double tiny = 1e-13 * sig[i0];
    if (sig[i0] > 0.0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) u[i0][r] = k[i0][r] / sig[i0];
    } else {
        u[i0][0] = 1.0; u[i0][1] = 0.0; u[i0][2] = 0.0;
    }
    if (sig[i1] > tiny) {
#pragma unroll
        for (int r = 0; r < 3; ++r) u[i1][r] = k[i1][r] / sig[i1];
    } else {                          // rank <= 1: any unit vector orthogonal to u0
        const double ax = fabs(u[i0][0]), ay = fabs(u[i0][1]), az = fabs(u[i0][2]);
        double e[3] = {0, 0, 0};
        e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
        const double d = e[0] * u[i0][0] + e[1] * u[i0][1] + e[2] * u[i0][2];
        double w[3] = {e[0] - d * u[i0][0], e[1] - d * u[i0][1], e[2] - d * u[i0][2]};
        const double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
        for (int r = 0; r < 3; ++r) u[i1][r] = w[r] / n;
    }
    if (sig[i2] > tiny) {
#pragma unroll
        for (int r = 0; r < 3; ++r) u[i2][r] = k[i2][r] / sig[i2];
    } else {                          // rank <= 2: complete the basis (either sign gives the same R, see Z below)
        u[i2][0] = u[i0][1] * u[i1][2] - u[i0][2] * u[i1][1];
        u[i2][1] = u[i0][2] * u[i1][0] - u[i0][0] * u[i1][2];
        u[i2][2] = u[i0][0] * u[i1][1] - u[i0][1] * u[i1][0];
    }
    const double dz = (det3(u) * det3(v) < 0.0) ? -1.0 : 1.0;     // sign(det(U V^T)); both factors are +-1
    double z[3] = {1.0, 1.0, 1.0};
    z[i2] = dz;
    const double scale = (z[0] * sig[0] + z[1] * sig[1] + z[2] * sig[2]) / var1;
    // R = V Z U^T  ->  R[r][c] = sum_i z_i v_i[r] u_i[c]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out.sR[r * 3 + c] = scale * (z[0] * v[0][r] * u[0][c] + z[1] * v[1][r] * u[1][c] + z[2] * v[2][r] * u[2][c]);
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out.t[r] = mu2[r] - (out.sR[r * 3 + 0] * mu1[0] + out.sR[r * 3 + 1] * mu1[1] + out.sR[r * 3 + 2] * mu1[2]);
}

// All four metrics of evaluate_pose for one sample (pose_3d_mvf_ex.py:317-333).  `pred`, `gt` in model units (cm);
// unit_scale = cm2mm.  out = {mpjpe*unit_scale (fp32 semantics), pa_mpjpe*unit_scale (fp64), pck_3d*100, auc_3d*100};
// s1_hat (optional, [J,3] double) receives the aligned prediction.
template <class Load>
EGR_HD void eval_pose_sample(Load pred, Load gt, int J, float unit_scale, float pck_thr, const float* auc_thr, int n_auc,
                             double out[4], double* s1_hat) {
    SimTransform T;
    similarity_transform(pred, gt, J, T);
    float sum_err = 0.f;
    double sum_pa = 0.0;
    int n_pck = 0, n_auc_hits = 0;
    for (int j = 0; j < J; ++j) {
        const float px = pred(j, 0), py = pred(j, 1), pz = pred(j, 2);
        const float gx = gt(j, 0), gy = gt(j, 1), gz = gt(j, 2);
        const float dx = px - gx, dy = py - gy, dz = pz - gz;
        sum_err += sqrtf(dx * dx + dy * dy + dz * dz);
        const double hx = T.sR[0] * px + T.sR[1] * py + T.sR[2] * pz + T.t[0];
        const double hy = T.sR[3] * px + T.sR[4] * py + T.sR[5] * pz + T.t[1];
        const double hz = T.sR[6] * px + T.sR[7] * py + T.sR[8] * pz + T.t[2];
        if (s1_hat) { s1_hat[j * 3 + 0] = hx; s1_hat[j * 3 + 1] = hy; s1_hat[j * 3 + 2] = hz; }
        const double ex = hx - (double)gx, ey = hy - (double)gy, ez = hz - (double)gz;
        sum_pa += sqrt(ex * ex + ey * ey + ez * ez);
        // PCK / AUC compare distances of the SCALED poses (pred * cm2mm, gt * cm2mm), fp32
        const float sx = gx * unit_scale - px * unit_scale, sy = gy * unit_scale - py * unit_scale,
                    sz = gz * unit_scale - pz * unit_scale;
        const float dist = sqrtf(sx * sx + sy * sy + sz * sz);
        n_pck += (dist <= pck_thr) ? 1 : 0;
        for (int t = 0; t < n_auc; ++t) n_auc_hits += (dist <= auc_thr[t]) ? 1 : 0;
    }
    out[0] = (double)((sum_err / (float)J) * unit_scale);
    out[1] = (sum_pa / (double)J) * (double)unit_scale;
    out[2] = (double)(((float)n_pck / (float)J) * 100.0f);
    out[3] = (n_auc > 0) ? (double)(((float)n_auc_hits / (float)(J * n_auc)) * 100.0f) : 0.0;
}

}  // namespace egr
