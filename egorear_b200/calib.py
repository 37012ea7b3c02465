"""Ego4View fisheye calibration constants (Scaramuzza world->camera polynomials).

Values are the calibration data the reference ships as JSON under
pose_estimation/utils/camera_calib_file/ego4view/camera_*.json (read by
pose_estimation/utils/camera_models.py:19-27).  They are embedded so the extension works without the
reference checkout; `load_calibration(dir)` reads the JSON files instead when a directory is given and
exists, exactly like `FishEyeCameraCalibratedModel.__init__`.
"""
import json
import os

CAMERA_NAMES = ("camera_front_left", "camera_front_right", "camera_back_left", "camera_back_right")

EGO4VIEW = {
    "camera_front_left": {
        "size": [872, 872],
        "image_center": [426.391953, 442.934625],
        "polynomialW2C": [460.338573, 334.319389, 66.308739, 52.536031, 28.965735, 11.448555, 3.534604, 0.641935],
    },
    "camera_front_right": {
        "size": [872, 872],
        "image_center": [441.06514, 423.106574],
        "polynomialW2C": [451.279153, 300.101709, 14.777962, 28.386413, 28.257463, 4.012532, -0.028834, 7.979098,
                          5.916757, 1.212304],
    },
    "camera_back_left": {
        "size": [872, 872],
        "image_center": [436.574771, 443.744199],
        "polynomialW2C": [455.217884, 304.117884, 30.117358, 37.914819, 25.500353, 7.907473, 3.490984, 2.700899,
                          0.72293],
    },
    "camera_back_right": {
        "size": [872, 872],
        "image_center": [437.88297, 438.370512],
        "polynomialW2C": [448.046216, 275.417817, -0.829026, 33.442463, 25.579666, -4.78561, 1.274275, 14.349067,
                          8.792641, 1.575194],
    },
}

# device->camera offsets of the synthetic rig, cm (utils/camera_models.py:29-40)
SYN_OFFSETS = {
    "camera_front_left": (6.0, 0.0, 0.0),
    "camera_front_right": (-6.0, 0.0, 0.0),
    "camera_back_left": (-6.0, 37.0, 0.0),
    "camera_back_right": (6.0, 37.0, 0.0),
}


def cameras_for(camera_model: str):
    """Camera list in the order `_reproject_3d_to_2d` concatenates them (egoposeformer_mvf_ex.py:340-382)."""
    if camera_model in ("ego4view_syn", "ego4view_rw"):
        return list(CAMERA_NAMES)
    if camera_model in ("ego4view_syn_stereo_front", "ego4view_rw_stereo_front"):
        return list(CAMERA_NAMES[:2])
    if camera_model in ("ego4view_syn_stereo_back", "ego4view_rw_stereo_back"):
        return list(CAMERA_NAMES[2:])
    raise ValueError("Unknown camera model !")  # same message as utils/camera_models.py:48


def load_calibration(calib_dir=None):
    """Return {camera_name: {"size", "image_center", "polynomialW2C"}}."""
    if calib_dir and os.path.isdir(calib_dir):
        out = {}
        for name in CAMERA_NAMES:
            path = os.path.join(calib_dir, name + ".json")
            if os.path.exists(path):
                with open(path) as f:
                    d = json.load(f)
                out[name] = {k: d[k] for k in ("size", "image_center", "polynomialW2C")}
            else:
                out[name] = EGO4VIEW[name]
        return out
    return {k: dict(v) for k, v in EGO4VIEW.items()}
