// Standalone operator entry points: the mmcv deformable-attention forward (the reference's only native op)
// and the fisheye reprojection.
#include "token_kernels.cuh"
#include "calib_data.h"

namespace egr {

// thread = (b, q, head, channel), channel fastest: the hd lanes of a head read hd contiguous floats per corner
__global__ void __launch_bounds__(256)
msda_forward_kernel(const float* __restrict__ value, int B, int H, int W, int nh, int hd, const float* __restrict__ loc,
                    const float* __restrict__ aw, int Q, int P, float* __restrict__ out) {
    const int64_t total = (int64_t)B * Q * nh * hd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % hd);
        int64_t r = i / hd;
        const int h = (int)(r % nh); r /= nh;
        const int q = (int)(r % Q);
        const int b = (int)(r / Q);
        const float* lp = loc + (((int64_t)b * Q + q) * nh + h) * P * 2;
        const float* ap = aw + (((int64_t)b * Q + q) * nh + h) * P;
        const float* vb = value + (int64_t)b * H * W * nh * hd + (int64_t)h * hd + c;
        float acc = 0.f;
        for (int p = 0; p < P; ++p) {
            const Corners cn = msda_corners(__ldg(lp + p * 2), __ldg(lp + p * 2 + 1), H, W);
            const float a = __ldg(ap + p);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                acc = fmaf(a * cn.w[k], __ldg(vb + (int64_t)cn.idx[k] * nh * hd), acc);
        }
        out[i] = acc;
    }
}

__global__ void reproject_kernel(float* __restrict__ pts3d, int B, int J, int V, int is_rw, int4 cam_id, CamCalib c0,
                                 CamCalib c1, CamCalib c2, CamCalib c3, const float* __restrict__ ctm,
                                 float* __restrict__ a2, uint8_t* __restrict__ av) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * J) return;
    const int b = i / J, j = i - b * J;
    float px = pts3d[i * 3 + 0], py = pts3d[i * 3 + 1], pz = pts3d[i * 3 + 2];
    const int ids[4] = {cam_id.x, cam_id.y, cam_id.z, cam_id.w};
    for (int v = 0; v < V; ++v) {
        const CamCalib& c = (v == 0) ? c0 : (v == 1) ? c1 : (v == 2) ? c2 : c3;
        float x, y, z;
        if (is_rw) {
            const float* M = ctm + ((int64_t)b * V + v) * 16;
            const float hx = px * 0.01f, hy = py * 0.01f, hz = pz * 0.01f;
            x = (M[0] * hx + M[1] * hy + M[2] * hz + M[3]) * 100.f;
            y = (M[4] * hx + M[5] * hy + M[6] * hz + M[7]) * 100.f;
            z = (M[8] * hx + M[9] * hy + M[10] * hz + M[11]) * 100.f;
        } else {
            float ox, oy;
            bool flip;
            syn_offset(ids[v], ox, oy, flip);
            if (flip) { px = -px; py = -py; }
            px = __fadd_rn(px, ox);
            py = __fadd_rn(py, oy);
            x = px; y = py; z = pz;
        }
        float u, vv;
        bool fov;
        fisheye_project(x, y, z, c, u, vv, fov);
        a2[(((int64_t)b * V + v) * J + j) * 2 + 0] = u;
        a2[(((int64_t)b * V + v) * J + j) * 2 + 1] = vv;
        av[((int64_t)b * V + v) * J + j] = fov ? 1 : 0;
    }
    if (!is_rw) { pts3d[i * 3 + 0] = px; pts3d[i * 3 + 1] = py; }   // the reference leaves the mutated tensor behind
}

CamCalib make_calib(int cam_id, const float* calib_host) {
    CamCalib c;
    const float* t = calib_host ? calib_host + cam_id * 16 : EGO4VIEW_CALIB[cam_id];
    c.cx = t[0]; c.cy = t[1]; c.size_h = t[2]; c.size_w = t[3];
    c.n_coef = (int)t[4];
    for (int i = 0; i < 11; ++i) c.coef[i] = t[5 + i];
    return c;
}

}  // namespace egr

using namespace egr;

extern "C" int egr_msda_forward(const float* value, int B, int H, int W, int nh, int hd, const float* loc,
                                const float* aw, int Q, int P, float* out, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(B >= 0 && H > 0 && W > 0 && nh > 0 && hd > 0 && Q > 0 && P > 0, EGR_ERR_INVALID, "msda: bad shape");
    if (B == 0) return EGR_OK;
    EGR_CHECK(value && loc && aw && out, EGR_ERR_INVALID, "msda: null pointer");
    const int64_t total = (int64_t)B * Q * nh * hd;
    const int64_t g = ceil_div64(total, 256);
    msda_forward_kernel<<<(int)(g < (int64_t)sm_count() * 32 ? g : (int64_t)sm_count() * 32), 256, 0, (cudaStream_t)stream>>>(
        value, B, H, W, nh, hd, loc, aw, Q, P, out);
    EGR_LAUNCHED();
    return EGR_OK;
}

extern "C" int egr_reproject_fisheye(float* pts3d, int B, int J, const int* cam_ids_host, int n_cams, int is_rw,
                                     const float* coord_trans_mat, const float* calib_host, float* anchors_2d,
                                     uint8_t* anchors_valid, void* stream) {
    if (int rc = require_device()) return rc;
    EGR_CHECK(n_cams >= 1 && n_cams <= 4 && cam_ids_host, EGR_ERR_INVALID, "reproject: n_cams=%d", n_cams);
    EGR_CHECK(!is_rw || coord_trans_mat, EGR_ERR_INVALID, "reproject: rw camera model needs coord_trans_mat");
    if (B == 0) return EGR_OK;
    int ids[4] = {0, 0, 0, 0};
    CamCalib c[4];
    for (int v = 0; v < 4; ++v) {
        if (v < n_cams) {
            EGR_CHECK(cam_ids_host[v] >= 0 && cam_ids_host[v] < 4, EGR_ERR_INVALID, "Unknown camera model !");
            ids[v] = cam_ids_host[v];
        }
        c[v] = make_calib(ids[v], calib_host);
    }
    reproject_kernel<<<ceil_div(B * J, 128), 128, 0, (cudaStream_t)stream>>>(
        pts3d, B, J, n_cams, is_rw, make_int4(ids[0], ids[1], ids[2], ids[3]), c[0], c[1], c[2], c[3], coord_trans_mat,
        anchors_2d, anchors_valid);
    EGR_LAUNCHED();
    return EGR_OK;
}
