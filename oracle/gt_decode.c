/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C CPU restatement of the two integer/bit-exact functions on the EgoRear
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the shipped path
 * (egorear_b200/) never does.
 *
 *   orc_generate_target  <- /root/reference/generate_heatmap.py:10-48  (generate_target)
 *   orc_get_max_preds    <- /root/reference/pose_estimation/utils/loss.py:122-142 (get_max_preds)
 *
 * Parity pinning: the reference ships no tests/golden vectors, so
 * this restatement is pinned against the reference itself: tests/golden/make_golden.py
 * imports /root/reference and stores its outputs under tests/golden/ (.npz files), and
 * tests/test_oracle.py checks this file against them bit for bit.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* fp32 Gaussian patch exp(-(dx^2+dy^2)/(2 sigma^2)) exactly as numpy evaluates it for sigma=1
 * (generate_heatmap.py:33-36: float32 arange, float32 arithmetic, np.exp on float32).
 * Index = dx^2+dy^2 in {0,1,2,4,5,8,9,10,13,18}; bit patterns recorded from numpy 2.3
 * and re-checked by tests/test_oracle.py against the live reference. */
static uint32_t sigma1_bits(int d2) {
    switch (d2) {
        case 0:  return 0x3F800000u; case 1:  return 0x3F1B4598u;
        case 2:  return 0x3EBC5AB1u; case 4:  return 0x3E0A9555u;
        case 5:  return 0x3DA81C2Fu; case 8:  return 0x3C960AAEu;
        case 9:  return 0x3C360282u; case 10: return 0x3BDCC9FEu;
        case 13: return 0x3AC50F0Cu; case 18: return 0x39016791u;
    }
    return 0u;
}

/* Fill patch[size*size] for integer tmp_size = 3*sigma (size = 2*tmp_size+1).
 * sigma == 1 uses the recorded numpy bit patterns; other sigmas evaluate the same
 * float32 expression with expf (may differ from numpy by 1 ulp: callers that need
 * bit-exactness for sigma != 1 pass numpy's own patch). */
void orc_gaussian_patch(double sigma, int size, float* patch) {
    int c = size / 2;
    for (int y = 0; y < size; ++y)
        for (int x = 0; x < size; ++x) {
            int d2 = (x - c) * (x - c) + (y - c) * (y - c);
            float v;
            if (sigma == 1.0) {
                uint32_t b = sigma1_bits(d2);
                memcpy(&v, &b, 4);
            } else {
                float num = -(float)d2;                       /* float32 numerator, as numpy */
                float den = (float)(2.0 * sigma * sigma);     /* python scalar -> float32 */
                v = expf(num / den);
            }
            patch[y * size + x] = v;
        }
}

/* generate_heatmap.py:10-48.  joints: [n_maps][J][2] float64 pixel coords (x,y) in image space,
 * out: [n_maps][J][hs][hs] float32.  patch: [size*size] float32 (size = 2*int(3*sigma)+1). */
void orc_generate_target(const double* joints, float* out, long n_maps, int J,
                         double image_size, int hs, double sigma, const float* patch) {
    const double tmp_size = sigma * 3.0;                       /* :18 */
    const int size = (int)(2.0 * tmp_size + 1.0);              /* :32 */
    const double feat_stride = image_size / (double)hs;        /* :21 */
    for (long m = 0; m < n_maps; ++m) {
        for (int j = 0; j < J; ++j) {
            float* tgt = out + ((size_t)m * J + j) * hs * hs;
            memset(tgt, 0, sizeof(float) * hs * hs);           /* np.zeros :13 */
            const double jx = joints[((size_t)m * J + j) * 2 + 0];
            const double jy = joints[((size_t)m * J + j) * 2 + 1];
            /* int() truncates toward zero :22-23 (C double->int cast does the same) */
            const int mu_x = (int)(jx / feat_stride + 0.5);
            const int mu_y = (int)(jy / feat_stride + 0.5);
            const int ul0 = (int)(mu_x - tmp_size), ul1 = (int)(mu_y - tmp_size);         /* :25 */
            const int br0 = (int)(mu_x + tmp_size + 1), br1 = (int)(mu_y + tmp_size + 1); /* :26 */
            if (ul0 >= hs || ul1 >= hs || br0 < 0 || br1 < 0) continue;                   /* :27-30 */
            const int gx0 = ul0 < 0 ? -ul0 : 0, gx1 = (br0 < hs ? br0 : hs) - ul0;        /* :38 */
            const int gy0 = ul1 < 0 ? -ul1 : 0, gy1 = (br1 < hs ? br1 : hs) - ul1;        /* :39 */
            const int ix0 = ul0 > 0 ? ul0 : 0;                                            /* :40 */
            const int iy0 = ul1 > 0 ? ul1 : 0;                                            /* :41 */
            for (int gy = gy0; gy < gy1; ++gy)
                for (int gx = gx0; gx < gx1; ++gx)
                    tgt[(iy0 + gy - gy0) * hs + (ix0 + gx - gx0)] = patch[gy * size + gx]; /* :45-46 */
        }
    }
}

/* utils/loss.py:122-142.  hm: [N][J][H][W] float32.
 * preds [N][J][2] (x,y), maxvals [N][J], valid [N][J] (uint8), idx [N][J] (int32 flat argmax, extra).
 * torch.max semantics: first index among equal maxima; a NaN is treated as the maximum
 * (first NaN wins), which torch.max on CPU/CUDA also does. */
void orc_get_max_preds(const float* hm, long N, int J, int H, int W, float threshold, int normalize,
                       float* preds, float* maxvals, uint8_t* valid, int32_t* idx_out) {
    const int HW = H * W;
    for (long n = 0; n < N * J; ++n) {
        const float* p = hm + (size_t)n * HW;
        float best = p[0];
        int bi = 0;
        if (!(best != best)) {
            for (int i = 1; i < HW; ++i) {
                float v = p[i];
                if (v != v) { best = v; bi = i; break; }
                if (v > best) { best = v; bi = i; }
            }
        }
        /* :131-134: idx -> float32, then `% W` and `// W` on floats.  torch's float remainder /
         * floor_divide are exact for integer-valued operands below 2^24 (fmod is exact and
         * (a - fmod(a,b)) / b is an exact integer), so integer arithmetic restates them exactly. */
        float x = (float)(bi % W);
        float y = (float)(bi / W);
        if (normalize) { x = x / (float)W; y = y / (float)H; }   /* :136-138 */
        preds[n * 2 + 0] = x;
        preds[n * 2 + 1] = y;
        maxvals[n] = best;
        valid[n] = (uint8_t)(best >= threshold);  /* :140 */
        if (idx_out) idx_out[n] = bi;
    }
}
