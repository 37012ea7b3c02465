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

// products that must round like the reference's separate torch ops (no FMA contraction: `gt*10 - pred*10` has to be
// exactly 0 when gt == pred, or PCK/AUC at threshold 0 change)
#if defined(__CUDA_ARCH__)
#define EGR_FMUL(a, b) __fmul_rn((a), (b))
#else
#define EGR_FMUL(a, b) ((a) * (b))
#endif

// MUFU-seeded helpers: the Jacobi rotations tolerate 2-ulp approximations (each sweep re-measures the angles; a common
// scale error of both rotated columns cancels when the frames are normalised in double), and a float rsqrt seed plus
// two Newton steps in double gives 1/sqrt(x) to ~1e-16 without the ~25-instruction IEEE double sqrt/div sequences.
#if defined(__CUDA_ARCH__)
#define EGR_RSQRTF(x) rsqrtf(x)
#define EGR_FDIV(a, b) __fdividef((a), (b))
#else
#define EGR_RSQRTF(x) (1.0f / sqrtf(x))
#define EGR_FDIV(a, b) ((a) / (b))
#endif

EGR_HD double rsqrt_d(double x) {            // x > 0 within float range
    double y = (double)EGR_RSQRTF((float)x);
    y = y * (1.5 - 0.5 * x * y * y);
    y = y * (1.5 - 0.5 * x * y * y);
    return y;
}
EGR_HD double sqrt_d(double x) { return ((float)x > 1e-30f) ? x * rsqrt_d(x) : sqrt(x); }

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix held column-wise, float32 like the reference's LAPACK sgesdd
// (numpy float32 input): on exit the columns of `a` are U*diag(sigma) and `v` holds V (K = U diag(sigma) V^T).
// Quadratically convergent; 3x3 needs 4-5 sweeps.  Only the two dominant singular pairs are used by the caller.
EGR_HD void jacobi_svd3(float a[3][3] /*[col][row]*/, float v[3][3] /*[col][row]*/) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) v[c][r] = (c == r) ? 1.f : 0.f;
    for (int sweep = 0; sweep < 10; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pair = 0; pair < 3; ++pair) {
            const int p = (pair == 2) ? 1 : 0;
            const int q = (pair == 0) ? 1 : 2;
            const float alpha = a[p][0] * a[p][0] + a[p][1] * a[p][1] + a[p][2] * a[p][2];
            const float beta = a[q][0] * a[q][0] + a[q][1] * a[q][1] + a[q][2] * a[q][2];
            const float gamma = a[p][0] * a[q][0] + a[p][1] * a[q][1] + a[p][2] * a[q][2];
            if (gamma * gamma > 1e-14f * alpha * beta) {          // |cos(angle)| > 1e-7
                rotated = true;
                // tan of the rotation angle, smaller root of t^2 + 2 zeta t - 1 = 0 with zeta = (beta-alpha)/(2 gamma):
                // t = sign(d) g / (|d| + sqrt(d^2 + g^2)),  d = beta - alpha, g = 2 gamma
                const float d = beta - alpha, g = 2.f * gamma;
                const float h2 = d * d + g * g;
                const float h = h2 * EGR_RSQRTF(h2);
                const float tt = EGR_FDIV((d >= 0.f) ? g : -g, fabsf(d) + h);
                const float cs = EGR_RSQRTF(1.f + tt * tt);
                const float sn = cs * tt;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float ap = a[p][r], aq = a[q][r];
                    a[p][r] = cs * ap - sn * aq;
                    a[q][r] = sn * ap + cs * aq;
                    const float vp = v[p][r], vq = v[q][r];
                    v[p][r] = cs * vp - sn * vq;
                    v[q][r] = sn * vp + cs * vq;
                }
            }
        }
        if (!rotated) break;
    }
}

// orthonormal pair from two (nearly orthogonal, unnormalised) float vectors, in double; e1 must be non-zero.
// ok2 = false: e2 carries no direction (rank <= 1) -> any unit vector orthogonal to e1.
EGR_HD void ortho_frame(const float x1[3], const float x2[3], bool ok2, double e1[3], double e2[3], double e3[3]) {
    double n = rsqrt_d((double)x1[0] * x1[0] + (double)x1[1] * x1[1] + (double)x1[2] * x1[2]);
#pragma unroll
    for (int r = 0; r < 3; ++r) e1[r] = (double)x1[r] * n;
    double w[3];
    if (ok2) {
        const double d = x2[0] * e1[0] + x2[1] * e1[1] + x2[2] * e1[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) w[r] = (double)x2[r] - d * e1[r];
    } else {
        const double ax = fabs(e1[0]), ay = fabs(e1[1]), az = fabs(e1[2]);
        const double b0 = (ax <= ay && ax <= az) ? 1.0 : 0.0;
        const double b1 = (b0 == 0.0 && ay <= az) ? 1.0 : 0.0;
        const double b2 = 1.0 - b0 - b1;
        const double d = b0 * e1[0] + b1 * e1[1] + b2 * e1[2];
        w[0] = b0 - d * e1[0]; w[1] = b1 - d * e1[1]; w[2] = b2 - d * e1[2];
    }
    n = rsqrt_d(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
    for (int r = 0; r < 3; ++r) e2[r] = w[r] * n;
    e3[0] = e1[1] * e2[2] - e1[2] * e2[1];
    e3[1] = e1[2] * e2[0] - e1[0] * e2[2];
    e3[2] = e1[0] * e2[1] - e1[1] * e2[0];
}

// Similarity transform (s, R, t) taking the J points `s1` onto `s2` (both [J,3] fp32):
//   mu, X = S - mu, var1 = sum X1^2, K = X1 X2^T (3x3), K = U S V^T,
//   reference: R = V diag(1, 1, sign det(U V^T)) U^T  (the sign lands on the SMALLEST singular value),
//              s = trace(R K) / var1, t = mu2 - s R mu1.
// With u3 := u1 x u2 and v3 := v1 x v2 (both frames right-handed) V [u1 u2 u3]^T IS that rotation for either sign of
// the determinant, so only the two dominant singular pairs of the float32 SVD are needed; they are re-orthonormalised
// in double, and s = trace(R K) / var1 is evaluated in double exactly as the reference writes it.
template <class Load>
EGR_HD void similarity_transform(Load s1, Load s2, int J, SimTransform& out) {
    double mu1[3] = {0, 0, 0}, mu2[3] = {0, 0, 0};
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) { mu1[c] += (double)s1(j, c); mu2[c] += (double)s2(j, c); }
    const double inv_j = 1.0 / (double)J;
#pragma unroll
    for (int c = 0; c < 3; ++c) { mu1[c] *= inv_j; mu2[c] *= inv_j; }
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
    // float32 SVD of K scaled to unit max-abs (no overflow / underflow in the squared norms)
    double kmax = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) kmax = fmax(kmax, fabs(k[c][r]));
    float a[3][3], v[3][3];
    const double inv = kmax > 0.0 ? 1.0 / kmax : 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) a[c][r] = (float)(k[c][r] * inv);
    jacobi_svd3(a, v);
    // bring the two largest columns to positions 0 and 1 (compare-exchange on registers, no indexed local arrays)
    float n0 = a[0][0] * a[0][0] + a[0][1] * a[0][1] + a[0][2] * a[0][2];
    float n1 = a[1][0] * a[1][0] + a[1][1] * a[1][1] + a[1][2] * a[1][2];
    float n2 = a[2][0] * a[2][0] + a[2][1] * a[2][1] + a[2][2] * a[2][2];
#define EGR_CSWAP(i, j, ni, nj)                                                                    \
    if (nj > ni) {                                                                                 \
        float tmp = ni; ni = nj; nj = tmp;                                                         \
        for (int r = 0; r < 3; ++r) {                                                              \
            tmp = a[i][r]; a[i][r] = a[j][r]; a[j][r] = tmp;                                       \
            tmp = v[i][r]; v[i][r] = v[j][r]; v[j][r] = tmp;                                       \
        }                                                                                          \
    }
    EGR_CSWAP(0, 1, n0, n1)
    EGR_CSWAP(1, 2, n1, n2)
    EGR_CSWAP(0, 1, n0, n1)
#undef EGR_CSWAP
    double R[9];
    if (n0 > 0.f) {
        const bool ok2 = n1 > 1e-12f * n0;           // sigma2 > 1e-6 sigma1; else K has rank <= 1 (collinear points)
        double u1[3], u2[3], u3[3], w1[3], w2[3], w3[3];
        ortho_frame(a[0], a[1], ok2, u1, u2, u3);
        ortho_frame(v[0], v[1], ok2, w1, w2, w3);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[r * 3 + c] = w1[r] * u1[c] + w2[r] * u2[c] + w3[r] * u3[c];   // R = V U^T
    } else {                                          // K == 0 (all points coincide): identity
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    double tr = 0.0;                                  // trace(R K) = sum_rc R[r][c] K[c][r]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) tr += R[r * 3 + c] * k[r][c];
    const double scale = tr / var1;
#pragma unroll
    for (int i = 0; i < 9; ++i) out.sR[i] = scale * R[i];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out.t[r] = mu2[r] - (out.sR[r * 3 + 0] * mu1[0] + out.sR[r * 3 + 1] * mu1[1] + out.sR[r * 3 + 2] * mu1[2]);
}

// All four metrics of evaluate_pose for one sample (pose_3d_mvf_ex.py:317-333).  `pred`, `gt` in model units (cm);
// unit_scale = cm2mm.  out = {mpjpe*unit_scale (fp32 semantics), pa_mpjpe*unit_scale (fp64), pck_3d*100, auc_3d*100};
// s1_hat (optional, [J,3] double) receives the aligned prediction.
template <class Load>
EGR_HD void eval_pose_sample(Load pred, Load gt, int J, float unit_scale, float pck_thr, const float* auc_thr, int n_auc,
                             bool auc_sorted, double out[4], double* s1_hat) {
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
        sum_pa += sqrt_d(ex * ex + ey * ey + ez * ez);
        // PCK / AUC compare distances of the SCALED poses (pred * cm2mm, gt * cm2mm), fp32
        const float sx = EGR_FMUL(gx, unit_scale) - EGR_FMUL(px, unit_scale), sy = EGR_FMUL(gy, unit_scale) - EGR_FMUL(py, unit_scale),
                    sz = EGR_FMUL(gz, unit_scale) - EGR_FMUL(pz, unit_scale);
        const float dist = sqrtf(sx * sx + sy * sy + sz * sz);
        n_pck += (dist <= pck_thr) ? 1 : 0;
        if (auc_sorted) {
            // ascending thresholds: #{t : dist <= thr[t]} = n - lower_bound(thr, dist); NaN compares false -> 0 hits
            // branch-free: pos = #{t : !(dist <= thr[t])} (a prefix for ascending thr), thresholds capped at 64
            int pos = 0;
#pragma unroll
            for (int step = 32; step > 0; step >>= 1) {
                const int q = pos + step;
                if (q <= n_auc && !(dist <= auc_thr[q - 1])) pos = q;
            }
            if (pos < n_auc && !(dist <= auc_thr[pos])) ++pos;         // n_auc == 64: the 64th element
            n_auc_hits += n_auc - pos;
        } else {
            for (int t = 0; t < n_auc; ++t) n_auc_hits += (dist <= auc_thr[t]) ? 1 : 0;
        }
    }
    out[0] = (double)((sum_err / (float)J) * unit_scale);
    out[1] = (sum_pa / (double)J) * (double)unit_scale;
    out[2] = (double)(((float)n_pck / (float)J) * 100.0f);
    out[3] = (n_auc > 0) ? (double)(((float)n_auc_hits / (float)(J * n_auc)) * 100.0f) : 0.0;
}

}  // namespace egr
