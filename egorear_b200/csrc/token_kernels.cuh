// Launchers of the two fused token kernels (see token_kernels.cu).
#pragma once
#include "token_common.cuh"

namespace egr {

// ---- mvfex refiner (HeatmapMVF.forward :655-711, jqa variant) ----
struct MvfTokenW {
    const float *hp2_T, *hp2_b;       // heatmap_proj.2        [256][256]
    const float *bfb_T, *bfb_b;       // fc_bfb                [512][256]
    const float *jq;                  // joint_query_embed     [J][256]
    const float *fcq_T, *fcq_b;       // fc_query.0            [256][256]
    LayerW layer;                     // transformer_layers.0
    const float *pn_w, *pn_b;         // post_norm.0
};

struct MvfTokenArgs {
    int B, V, J, H, W;
    int r0, G;                        // group g (blockIdx.y) uses refiner weights r0 + g
    const float* q1;                  // [G][B][J][256] relu(heatmap_proj.0(heatmap)) fp32
    const float* bfb;                 // stride-32 backbone map of group g, frame b: bfb + g*bfb_gs + b*bfb_bs, [512][bfb_hw]
    int64_t bfb_bs, bfb_gs;
    int bfb_hw;
    const float* anchors;             // [B][V][J][2]
    const uint8_t* valid;             // [B][V][J]
    const void* X;                    // [V][B][H*W][128] channels-last features (float | bf16)
    void* xT;                         // out [G][B][256][16]: post_norm tokens transposed (pos-major, joint-minor, col 15 = 0)
    const MvfTokenW* w;               // device array [V]
};
int launch_mvf_tokens(const MvfTokenArgs& a, int act_bf16, cudaStream_t st);

// ---- pose3d lifting (EgoPoseFormerPose3D._forward_transformer :384-420 + reprojection :340-382) ----
struct PoseTokenW {
    const float *g0_T, *g0_b;         // query_gen_mlp.0   [4][128]
    const float *g2_T, *g2_b;         // query_gen_mlp.2   [128][128]
    const float *g4_T, *g4_b;         // query_gen_mlp.4   [128][128]
    LayerW layer[4];
    const float *pn_w[4], *pn_b[4];   // post_norm.i
    const float *r0_T[4], *r0_b[4];   // reg_mlp.i.0       [128][128]
    const float *r2_T[4], *r2_b[4];   // reg_mlp.i.2       [128][3] (transposed: [k][3])
};

struct CamCalib {            // one camera, host/device POD
    float cx, cy, size_h, size_w;
    int n_coef;
    float coef[11];
};

struct PoseTokenArgs {
    int B, V, J, H, W, L;
    int is_rw;
    int cam_id[4];                    // 0 FL 1 FR 2 BL 3 BR, in concat order
    CamCalib cam[4];                  // indexed by view slot
    const float* ctm;                 // [B][V][4][4] (rw) or null
    const float* mlp_pred;            // [B][J][3] proposal (preds[0])
    float* preds;                     // [L+1][B][J][3]; rows 1..L written here
    const void* X;                    // [V][B][H*W][128] channels-last frame_feats (float | bf16)
    const PoseTokenW* w;              // device pointer (single struct)
    float* dbg_anchors;               // optional [B][V][J][2]
    uint8_t* dbg_valid;               // optional [B][V][J]
};
int launch_pose_tokens(const PoseTokenArgs& a, int act_bf16, cudaStream_t st);

// FishEyeCameraCalibratedModel.world2camera_pytorch (utils/camera_models.py:70-104) for one camera-relative point.
// Elementwise fp32 op order of the reference is kept (no FMA contraction): norm = sqrt(x*x + y*y),
// theta = atan(-z / norm), rho = sum_i a_i * theta**i (powers, summed from i = 0), u = (x / norm * rho + cx) / w.
__device__ __forceinline__ void fisheye_project(float x, float y, float z, const CamCalib& c, float& u, float& v,
                                                bool& in_fov) {
    const float norm = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
    const float theta = atanf(__fdiv_rn(-z, norm));
    float rho = 0.f;
    for (int i = 0; i < c.n_coef; ++i) {
        float t;
        if (i == 0) t = 1.f;
        else if (i == 1) t = theta;
        else if (i == 2) t = __fmul_rn(theta, theta);
        else if (i == 3) t = __fmul_rn(__fmul_rn(theta, theta), theta);
        else t = powf(theta, (float)i);
        rho = __fadd_rn(rho, __fmul_rn(c.coef[i], t));
    }
    u = __fdiv_rn(__fadd_rn(__fmul_rn(__fdiv_rn(x, norm), rho), c.cx), c.size_w);
    v = __fdiv_rn(__fadd_rn(__fmul_rn(__fdiv_rn(y, norm), rho), c.cy), c.size_h);
    in_fov = (u > 0.f) && (v > 0.f) && (u < 1.f) && (v < 1.f);
    u = (u != u) ? u : fminf(fmaxf(u, 0.f), 1.f);   // clamp keeps NaN like torch.clamp
    v = (v != v) ? v : fminf(fmaxf(v, 0.f), 1.f);
}

// device->camera transform of one view slot.  syn: cumulative IN-PLACE flips/offsets exactly like
// get_camera_relative_pts3d (utils/camera_models.py:53-68); rw: (M · [p*0.01; 1]) * 100, out of place.
__device__ __forceinline__ void syn_offset(int cam_id, float& ox, float& oy, bool& flip) {
    // utils/camera_models.py:29-40, :57-61
    ox = (cam_id == 0 || cam_id == 3) ? 6.f : -6.f;
    oy = (cam_id >= 2) ? 37.f : 0.f;
    flip = cam_id >= 2;
}

}  // namespace egr
