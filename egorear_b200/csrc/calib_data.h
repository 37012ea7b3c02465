// Ego4View calibration table for the C side: per camera {cx, cy, size_h, size_w, n_coef, a0..a10}.
// Same numbers as egorear_b200/calib.py (source: the reference's camera_calib_file/ego4view/*.json).
#pragma once
static const float EGO4VIEW_CALIB[4][16] = {
    {426.391953f, 442.934625f, 872.0f, 872.0f, 8.0f, 460.338573f, 334.319389f, 66.308739f, 52.536031f, 28.965735f, 11.448555f, 3.534604f, 0.641935f, 0.0f, 0.0f, 0.0f},  // camera_front_left
    {441.06514f, 423.106574f, 872.0f, 872.0f, 10.0f, 451.279153f, 300.101709f, 14.777962f, 28.386413f, 28.257463f, 4.012532f, -0.028834f, 7.979098f, 5.916757f, 1.212304f, 0.0f},  // camera_front_right
    {436.574771f, 443.744199f, 872.0f, 872.0f, 9.0f, 455.217884f, 304.117884f, 30.117358f, 37.914819f, 25.500353f, 7.907473f, 3.490984f, 2.700899f, 0.72293f, 0.0f, 0.0f},  // camera_back_left
    {437.88297f, 438.370512f, 872.0f, 872.0f, 10.0f, 448.046216f, 275.417817f, -0.829026f, 33.442463f, 25.579666f, -4.78561f, 1.274275f, 14.349067f, 8.792641f, 1.575194f, 0.0f},  // camera_back_right
};
