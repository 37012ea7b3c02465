#!/usr/bin/env python
"""bench.py — EgoRear hot path on B200: 4-view frames/s of the mvfex-n1_jqa heatmap + pose3d forward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W] [--batch B]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step is one pass of the hot path over one batch of synthetic backbone features per GPU
(configs[1] of BASELINE.json: ego4view_syn_heatmap_mvfex-n1_jqa, batch 64, + the pose3d lifting of the metric):
H1 D1 Q1 M1 F1 A1-3 T1 R1 H2 -> decode -> P1-P4 -> pack -> (N>1) one NCCL all-gather of 672 B/frame.
Frames shard across ranks with no other collective ("scaling": "weak", fixed 64 frames per GPU).

Rank 0 prints ONE JSON line: value = whole-job frames/s with inputs resident in HBM; e2e = the same through the
public API with pinned-host inputs (H2D + D2H inside the timed region); roofline = dominant kernel (CUDA events,
stage profiler of the library) against MEASURED_PEAKS.json; cpu_baseline = the CPU implementation on the host cores
(the reference's own modules from /root/reference or its file copy oracle/_ref, else the oracle port) on a bounded sample;
parity = the oracle on 8 sampled frames of the timed batch (untimed); sustained = the same step back to back for >= 3 s.

Other workloads (parity-test configs, not the headline): --workload generate_target | decode | pose3d | mvfex | rw_e2e,
and the widened rows of SURVEY §8f: eval_heatmap | eval_pose (eval-time metrics) and preprocess (PIL-exact resize +
normalise); each prints the same JSON line with its own roofline, cpu_baseline and e2e.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "4-view frames/sec (heatmap+pose3d fwd)"
# the tensor-core precision that is INSIDE the fp32 parity bound (<= 1e-3 relative): fp16 operands, split 1x1 weights,
# 3x-TF32 / fp16-pair token Linears, activation pairs on the refine path (include/egorear_b200.h EGR_PREC_FP16);
# "bf16" is the looser-bound mode of round 1
DEFAULT_PRECISION = "fp16"
UNIT = "frames/s"

# algorithmic FLOPs per 4-view frame of every dense stage (2*M*N*K, with the 1x1 convs that follow a bilinear x2
# commuted in front of it), and algorithmic HBM bytes per frame in bf16 mode
STAGE_FLOPS = {
    "H1a": 4 * 2 * 4096 * 128 * 128, "H1b": 4 * 2 * 1024 * 256 * 1152, "H1c": 4 * 2 * 1024 * 256 * 256,
    "H1d": 4 * 2 * 1024 * 128 * 256, "Q1a": 4 * 2 * 15 * 256 * 4096, "F1a": 4 * 2 * 4096 * 256 * 128,
    "F1b": 4 * 2 * 1024 * 512 * 2304, "F1c": 4 * 2 * 1024 * 128 * 512, "R1a": 4 * 2 * 1024 * 128 * 128,
    "R1b": 4 * 2 * 1024 * 128 * 128, "H2a": 4 * 2 * 1024 * 256 * 1152, "H2b": 4 * 2 * 1024 * 256 * 256,
    "H2c": 4 * 2 * 1024 * 128 * 256,
    "P2a": 4 * 2 * 4096 * 64 * 128, "P2b": 4 * 2 * 1024 * 128 * 576, "P2c": 4 * 2 * 256 * 64 * 128,
    "P2d": 4 * 2 * 64 * 128 * 576, "P2mlp0": 2 * 2048 * 32768,
    # backbone engine, per 4-view frame: ResNet18 stages + EfficientFPN (26.7 GFLOP in total)
    "B_stem": 4 * 2 * 16384 * 64 * 147,
    "B_layer1": 4 * (4 * 2 * 4096 * 64 * 576 + 2 * 4096 * 128 * 64),
    "B_layer2": 4 * (2 * 1024 * 128 * 576 + 2 * 1024 * 128 * 64 + 3 * 2 * 1024 * 128 * 1152 + 2 * 1024 * 128 * 128),
    "B_layer3": 4 * (2 * 256 * 256 * 1152 + 2 * 256 * 256 * 128 + 3 * 2 * 256 * 256 * 2304 + 2 * 256 * 128 * 256),
    "B_layer4": 4 * (2 * 64 * 512 * 2304 + 2 * 64 * 512 * 256 + 3 * 2 * 64 * 512 * 4608 + 2 * 64 * 128 * 512),
    "B_fpn": 4 * sum(2 * (r // 2) ** 2 * 128 * 128 + 2 * r * r * 128 * 128 + 2 * r * r * 128 * 1152 for r in (16, 32, 64)),
}
# algorithmic HBM bytes per 4-view frame of every stage (act = bytes per activation element: 2 in bf16 mode; the pose3d
# proposal branch P2* keeps fp32/TF32 activations; weights are counted once per batch for the one stage where they
# dominate, P2mlp0).  `exp` = 1 when the forward also exports the TF32 channels-last refined features, 2 when in addition the NCHW fp32 refined
# features are not materialised (chained model; with the fp16 proposal branch the fp16 copy is then the only one).
STAGE_BYTES = {
    "stage_nhwc": lambda act, exp, B: 4 * 4096 * 128 * (4 + act),
    "H1a": lambda act, exp, B: 4 * 4096 * 128 * act * 2,
    "H1b": lambda act, exp, B: 4 * (4096 * 128 + 1024 * 256) * act,
    "H1c": lambda act, exp, B: 4 * 1024 * 256 * act * 2,
    "H1d": lambda act, exp, B: 4 * 1024 * (256 + 128) * act,
    "H1tail": lambda act, exp, B: 4 * (1024 * 128 * act + 15 * 4096 * (4 + act)),
    "D1": lambda act, exp, B: 4 * 15 * 4096 * 4,
    "Q1a": lambda act, exp, B: 4 * 15 * 4096 * act + 4 * 256 * 4096 * act // B,
    "F1a": lambda act, exp, B: 4 * 4096 * (128 + 256) * act,
    "F1b": lambda act, exp, B: 4 * (4096 * 256 + 1024 * 512) * act,
    "F1c": lambda act, exp, B: 4 * 1024 * (512 + 128) * act,
    "R1a": lambda act, exp, B: 4 * 1024 * 256 * act,
    "R1b": lambda act, exp, B: 4 * 1024 * 256 * act,
    "R1tail": lambda act, exp, B: 4 * (1024 * 128 * act + 4096 * 128 * (
        (0 if exp == 2 else 4) + (0 if (exp == 2 and P2ACT[0] == 2) else act) + (P2ACT[0] if exp else 0))),
    "H2a": lambda act, exp, B: 4 * (4096 * 128 + 1024 * 256) * act,
    "H2b": lambda act, exp, B: 4 * 1024 * 512 * act,
    "H2c": lambda act, exp, B: 4 * 1024 * 384 * act,
    "H2tail": lambda act, exp, B: 4 * (1024 * 128 * act + 15 * 4096 * 4),
    "P_stage_nhwc": lambda act, exp, B: 0 if exp else 4 * 4096 * 128 * (4 + P2ACT[0]) + 4 * 4096 * 128 * (4 + act),
    "P2a": lambda act, exp, B: 4 * 4096 * (128 + 64) * P2ACT[0],
    "P2b": lambda act, exp, B: 4 * (4096 * 64 + 1024 * 128) * P2ACT[0],
    "P2pool": lambda act, exp, B: 4 * (1024 + 256) * 128 * P2ACT[0],
    "P2c": lambda act, exp, B: 4 * 256 * (128 + 64) * P2ACT[0],
    "P2d": lambda act, exp, B: 4 * (256 * 64 + 64 * 128) * P2ACT[0],
    "P2mlp0": lambda act, exp, B: 2048 * 32768 * P2ACT[0] // B + 4 * 64 * 128 * P2ACT[0],
    # backbone stages: what their kernels must move at least (every conv reads its input and writes its output once, residual
    # branches are read once; the stem reads the fp32 image and writes the pooled map)
    "B_stem": lambda act, exp, B: 4 * (3 * 256 * 256 * 4 + 64 * 64 * 64 * act),
    "B_layer1": lambda act, exp, B: 4 * 4096 * 64 * act * 10 + 4 * 4096 * (64 + 128) * act,
    "B_layer2": lambda act, exp, B: 4 * (4096 * 64 * act * 2 + 1024 * 128 * act * 10) + 4 * 1024 * 256 * act,
    "B_layer3": lambda act, exp, B: 4 * (1024 * 128 * act * 2 + 256 * 256 * act * 10) + 4 * 256 * 384 * act,
    "B_layer4": lambda act, exp, B: 4 * (256 * 256 * act * 2 + 64 * 512 * act * 10) + 4 * 64 * 640 * act,
    "B_fpn": lambda act, exp, B: 4 * sum((r // 2) ** 2 * 256 + r * r * 128 * 4 for r in (16, 32, 64)) * act,
}
P2ACT = [4]      # bytes per element of the pose3d proposal branch: 4 (fp32 / TF32) or 2 (fp16, the bf16-mode default); set in main()
TF32_STAGES = ("P2a", "P2b", "P2c", "P2d", "P2mlp0")     # kind::tf32: half the bf16 tensor rate (nominal ratio)


def stage_roofline(name, ms, B, act, exp, peaks):
    """roofline of one stage: the binding resource is whichever of (FLOPs / tensor peak, bytes / HBM peak) takes longer"""
    fl = STAGE_FLOPS.get(name, 0) * B
    by = STAGE_BYTES[name](act, exp, B) * B if name in STAGE_BYTES else 0
    if not fl and not by:
        return None
    # stage times are CUDA-event intervals of a few-step pass, i.e. kernels timed (almost) alone: the tensor denominator is the
    # BURST cuBLAS figure; the seconds-long `sustained` block of the line is judged against the sustained one
    tf_peak = peaks["tf_burst"] * (0.5 if (name in TF32_STAGES and P2ACT[0] == 4) else 1.0)
    t_tensor = fl / (tf_peak * 1e12) if fl else 0.0
    t_hbm = by / (peaks["hbm"] * 1e9) if by else 0.0
    t_s = ms / 1e3
    if t_tensor >= t_hbm:
        ach = fl / t_s / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak}
    ach = by / t_s / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"]}


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage kernels, from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic from the .ncu-rep)"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
        return {k: int(v["dram_read_bytes"]) + int(v["dram_write_bytes"]) for k, v in d.get("stages", {}).items()}, d.get("source")
    except Exception:
        return {}, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")   # B200_PROFILING.md fallback


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.02)
        except Exception as e:      # NVML missing: report that instead of dying
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------------
# CPU implementation of the path (cpu_baseline leg and --impl reference arm)
# --------------------------------------------------------------------------------------------------------
class CpuPath:
    """the reference's own modules when a checkout (build container) or its file copy oracle/_ref (GPU box) is present
    (kind 'reference'), else the oracle port (kind 'port')"""

    def __init__(self, camera_model="ego4view_syn"):
        import torch
        from egorear_b200 import calib, synth
        from egorear_b200.configs import heatmap_mvfex_cfg, pose3d_cfg
        from egorear_b200.modules import EgoPoseFormerHeatmapMVFEX, EgoPoseFormerPose3D
        from oracle import model_ref, ref_import     # cpu_baseline / reference arm: the one place bench.py runs oracle/
        self.torch, self.model_ref = torch, model_ref
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.calib = calib.load_calibration(None)
        self.cam = camera_model
        self.kind = "port"
        self.how = "oracle/model_ref.py (torch fp32 restatement)"
        hm = synth.fill_state_dict(EgoPoseFormerHeatmapMVFEX(**heatmap_mvfex_cfg(4, camera_model), precision="fp32", build_backbone=False))
        p3 = synth.fill_state_dict(EgoPoseFormerPose3D(**pose3d_cfg(4, camera_model), precision="fp32"))
        self.sd_h, self.sd_p = hm.state_dict(), p3.state_dict()
        self.ref_h = self.ref_p = None
        if ref_import.available():
            try:
                import copy
                ref_import.FORCE_CPU = True          # the GPU box has CUDA: keep the camera model's tensors on the host
                cls = ref_import.import_estimators()
                rw = camera_model.startswith("ego4view_rw")
                cfg = ref_import.load_model_cfg("ego4view_%s_heatmap_mvfex-n1_jqa.yaml" % ("rw" if rw else "syn"))
                self.ref_h = cls["EgoPoseFormerHeatmapMVFEX"](**copy.deepcopy(cfg)).eval()
                self.ref_h.load_state_dict({**self.ref_h.state_dict(), **self.sd_h}, strict=True)
                c3 = ref_import.load_model_cfg("ego4view_%s_pose3d.yaml" % ("rw" if rw else "syn"))["pose3d_cfg"]
                c3.update(dict(num_views=4, image_size=[256, 256], use_pred_heatmap_init=True, camera_model=camera_model))
                self.ref_p = cls["EgoPoseFormerPose3D"](**c3).eval()
                self.ref_p.load_state_dict(self.sd_p, strict=True)
                fns = ref_import.import_functions()
                self.get_max_preds = fns["get_max_preds"]
                self.kind = "reference"
                self.how = "the reference's own modules (%s; mmcv op -> grid_sample shim)" % (
                    "file copy oracle/_ref" if ref_import.is_copy() else "checkout")
            except Exception as e:      # fall back to the port, say why
                sys.stderr.write("bench: live reference unusable (%s); timing the oracle port\n" % e)
                self.ref_h = self.ref_p = None

    def mvfex(self, feat, bfb):
        torch = self.torch
        with torch.no_grad():
            if self.ref_h is not None:
                self.ref_h.forward_heatmap_feat_estimation = lambda img: (feat, [None, None, None, bfb])
                lh, lf = self.ref_h(torch.zeros(feat.shape[0], 4, 3, 1, 1))
                B, V, J, H, W = lh[-1].shape
                self.get_max_preds(lh[-1].view(B * V, J, H, W), 0.5, False)
                return lh, lf
            lh, lf, _, _ = self.model_ref.mvfex_hot_path(self.sd_h, feat, bfb)
            B, V, J, H, W = lh[-1].shape
            self.model_ref.get_max_preds(lh[-1].view(B * V, J, H, W), 0.5, False)
            return lh, lf

    def pose3d(self, fi, ff, ctm=None):
        with self.torch.no_grad():
            if self.ref_p is not None:
                return self.ref_p(fi, ff, None, ctm)[-1]
            return self.model_ref.pose3d_forward(self.sd_p, fi, ff, self.calib, self.cam, ctm)[-1]

    def step(self, feat, bfb, ctm=None):
        lh, lf = self.mvfex(feat, bfb)
        return self.pose3d(lf[0], lf[-1], ctm)


def _timed(fn, unit_per_call, budget_s, max_units):
    fn()                                         # warm-up
    n, t0 = 0, time.time()
    while True:
        fn()
        n += unit_per_call
        if time.time() - t0 > budget_s or n >= max_units:
            break
    return n, time.time() - t0


def cpu_baseline(workload="mvfex_pose3d", budget_s=12.0):
    """the CPU implementation of the workload on the host cores, bounded sample; batch sizes per BASELINE.md section 4
    (mvfex B=4, pose3d B=32): the reference's modules when a checkout or its file copy oracle/_ref is present, else the port"""
    import torch
    from egorear_b200 import synth
    if workload in ("generate_target", "decode"):
        return cpu_baseline_gt_decode(workload, min(budget_s, 8.0))
    cam = "ego4view_rw" if workload == "rw_e2e" else "ego4view_syn"
    cp = CpuPath(cam)
    if workload == "pose3d":
        fps = 32
        g = torch.Generator().manual_seed(0)
        fi = torch.relu(torch.randn((fps, 4, 128, 64, 64), generator=g))
        ff = torch.relu(torch.randn((fps, 4, 128, 64, 64), generator=g))
        fn = lambda: cp.pose3d(fi, ff)
        what = "pose3d lifting"
    else:
        fps = 4
        feat, bfb = synth.synth_features(fps, 4, seed=0)
        ctm = synth.synth_coord_trans_mat(fps, seed=0) if cam == "ego4view_rw" else None
        if workload == "mvfex":
            fn = lambda: cp.mvfex(feat, bfb)
            what = "mvfex hot path + decode"
        else:
            fn = lambda: cp.step(feat, bfb, ctm)
            what = "mvfex hot path + decode + pose3d lifting" + (" (backbone excluded)" if workload == "rw_e2e" else "")
    n, dt = _timed(fn, fps, budget_s, 512)
    return {"value": n / dt, "unit": UNIT, "cores": cp.cores, "kind": cp.kind,
            "sample": "%d frames (steps of %d): %s, %s, torch fp32 on %d host threads, %.1f s" % (n, fps, what, cp.how, cp.cores, dt)}


def cpu_baseline_gt_decode(workload, budget_s=8.0):
    """generate_target as shipped (numpy, one call per camera as in the main loop of generate_heatmap.py:58-68) on one
    core when the reference is present, else the C oracle; get_max_preds through torch on all cores"""
    import numpy as np
    import torch
    from egorear_b200 import synth
    from oracle import model_ref, ref_import
    if workload == "decode":
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n = 64
        hm = torch.randn((n * 4, 15, 64, 64), generator=torch.Generator().manual_seed(0))
        kind, fn = "port", (lambda: model_ref.get_max_preds(hm, 0.5, True))
        how = "oracle/model_ref.get_max_preds (the reference's torch lines restated)"
        if ref_import.available():
            try:
                ref_fn = ref_import.import_functions()["get_max_preds"]
                kind, fn, how = "reference", (lambda: ref_fn(hm, 0.5, True)), "utils/loss.py:get_max_preds"
            except Exception:
                pass
        done, dt = _timed(fn, n, budget_s, 1 << 16)
        return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": "%d frames (steps of %d) x 4 views x 15 joints, %s, torch on %d threads, %.1f s" % (done, n, how, cores, dt)}
    n = 64
    kp = synth.synth_keypoints(n, 4, 16, seed=0)
    kind, how, fn = "port", "oracle/gt_decode.c (C restatement)", None
    if ref_import.available():
        try:
            gen = ref_import.import_functions()["generate_target"]
            fn = lambda: [gen(kp[f, v], 872, 64, 16, 1.0) for f in range(n) for v in range(4)]
            kind, how = "reference", "generate_heatmap.generate_target (numpy, one call per camera as in its main loop)"
        except Exception:
            fn = None
    if fn is None:
        import ctypes
        import subprocess
        so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        orc = ctypes.CDLL(so)
        orc.orc_gaussian_patch.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
        orc.orc_generate_target.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
        patch = np.zeros((7, 7), np.float32)
        orc.orc_gaussian_patch(1.0, 7, patch.ctypes.data)
        out = np.empty((n, 4, 16, 64, 64), np.float32)
        kpc = np.ascontiguousarray(kp)
        fn = lambda: orc.orc_generate_target(kpc.ctypes.data, out.ctypes.data, n * 4, 16, 872.0, 64, 1.0, patch.ctypes.data)
    done, dt = _timed(fn, n, budget_s, 1 << 16)
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d frames (steps of %d) x 4 views x 16 joints, %s, one core, %.1f s" % (done, n, how, dt)}


def cpu_baseline_eval(workload, budget_s=8.0):
    """the oracle's numpy restatement of the wrappers' metric methods on one host core (the reference itself runs them
    as per-sample Python loops over device tensors)"""
    from egorear_b200 import synth
    from oracle import metrics_ref as mr
    if workload == "eval_pose":
        n = 2048
        a, b = synth.synth_eval_poses(n, 16, seed=0)
        fn = lambda: mr.evaluate_pose(a, b)
    else:
        n = 16
        a, b = synth.synth_eval_heatmaps(n, 4, 15, seed=0)
        fn = lambda: mr.evaluate_heatmap(a, b)
    fn()
    done, t0 = 0, time.time()
    while time.time() - t0 < budget_s and done < 64 * n:
        fn()
        done += n
    dt = time.time() - t0
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d units (steps of %d) through oracle/metrics_ref.py (numpy), %.1f s" % (done, n, dt)}


def cpu_baseline_preprocess(budget_s=8.0):
    """PIL + torchvision exactly as the reference's datasets call them (kind "reference") when importable, else the
    oracle's numpy restatement; one host core, 4 views per frame."""
    from egorear_b200 import synth
    imgs = synth.synth_images(4, 872, 872, seed=0)
    try:
        from PIL import Image
        from torchvision import transforms
        tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
        pil = [Image.fromarray(i) for i in imgs]
        fn = lambda: [tf(p.convert("RGB").resize([256, 256], Image.BICUBIC)).float().numpy() for p in pil]
        kind, how = "reference", "PIL %s resize + torchvision ToTensor/Normalize" % __import__("PIL").__version__
    except ImportError:
        from oracle import preprocess_ref as pr
        fn = lambda: [pr.preprocess(i)[0] for i in imgs]
        kind, how = "port", "oracle/preprocess_ref.py (numpy)"
    fn()
    done, t0 = 0, time.time()
    while time.time() - t0 < budget_s and done < 256:
        fn()
        done += 1
    dt = time.time() - t0
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d frames x 4 views of 872x872 (decoded, in memory), %s, %.1f s" % (done, how, dt)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    from egorear_b200 import synth
    cp = CpuPath()
    fps = 4                                      # the CPU arm runs the mvfex path at batch 4
    feat, bfb = synth.synth_features(fps, 4, seed=0)
    for _ in range(max(1, min(args.warmup, 2))):
        cp.step(feat, bfb)
    steps = max(1, min(args.steps, 8))           # bounded: each step is a 4-frame sample of the workload
    t0 = time.time()
    for _ in range(steps):
        cp.step(feat, bfb)
    dt = time.time() - t0
    v = steps * fps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample_frames_per_step": fps, "note": "CPU implementation on host cores"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cp.cores, "kind": cp.kind,
                             "sample": "%d steps x %d frames, %s, torch fp32, %d host threads" % (steps, fps, cp.how, cp.cores)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def workload_name(args):
    return {"mvfex_pose3d": "ego4view_syn_heatmap_mvfex-n1_jqa + ego4view_syn_pose3d forward (hot path, backbone features as inputs), "
                            "4 views, batch %d/GPU" % args.batch,
            "mvfex": "ego4view_syn_heatmap_mvfex-n1_jqa hot path, batch %d/GPU" % args.batch,
            "pose3d": "ego4view_syn_pose3d lifting, batch %d/GPU" % args.batch,
            "generate_target": "generate_target sweep, 4 views x 16 joints, %d frames/step/GPU" % args.batch,
            "decode": "get_max_preds, 4 views x 15 joints, %d frames/step/GPU" % args.batch,
            "eval_heatmap": "eval-time heatmap metrics (wrapper `evaluate`: L1, positive L1, MSE, arg-max MSE), 4 views x 15 joints, "
                            "%d frames/step/GPU" % args.batch,
            "preprocess": "dataset preprocessing: PIL bicubic 872x872 -> 256x256 + ToTensor + Normalize, 4 views, %d frames/step/GPU" % args.batch,
            "eval_pose": "eval-time pose metrics (wrapper `evaluate_pose`: MPJPE, PA-MPJPE, PCK, AUC), 16 joints, "
                         "%d poses/step/GPU" % args.batch,
            "rw_e2e": "ego4view_rw_heatmap_mvfex-n1_jqa + ego4view_rw_pose3d from images (ResNet18 + EfficientFPN on the backbone engine + hot path), "
                      "batch %d/GPU" % args.batch}[args.workload]


# --------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _quiet_stdout():
    """Route fd 1 to stderr for the run (NCCL prints its version banner to stdout); the JSON line goes to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mvfex_pose3d", choices=["mvfex_pose3d", "mvfex", "pose3d", "generate_target", "decode", "rw_e2e", "eval_heatmap", "eval_pose", "preprocess"])
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("EGR_PRECISION", DEFAULT_PRECISION), choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--sustain-s", type=float, default=3.0, help="seconds of the extra back-to-back loop behind the `sustained` block (0 = off)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity block (oracle on a sample of the batch)")
    ap.add_argument("--lanes", type=int, default=3, help="streams alternated by the throughput loop (1 = one stream)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=int (egr_set_option), e.g. pdl=0")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {"mvfex_pose3d": 64, "mvfex": 64, "pose3d": 1024, "generate_target": 8192, "decode": 8192,
                      "rw_e2e": 512, "eval_heatmap": 2048, "eval_pose": 1 << 20, "preprocess": 256}[args.workload]
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    from egorear_b200 import _lib, dist as egd, ops, synth
    from egorear_b200.pipeline import HotPathPipeline

    rank, local_rank, world = egd.init()
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    for kv in args.opt:
        k, v = kv.split("=")
        _lib.check(lib.egr_set_option(k.encode(), int(v)))
    peaks = load_peaks()
    B = args.batch
    act = 2 if args.precision in ("bf16", "fp16") else 4
    ncu_traffic, ncu_source = load_ncu_traffic()
    parity_fn = None
    e2e_fn = None

    if args.workload in ("mvfex_pose3d", "mvfex", "pose3d"):
        # mvfex_pose3d = the chained model (EgoPoseFormerMVFEX.forward returns poses + heatmaps): refined features internal
        pipe = HotPathPipeline(4, "ego4view_syn", args.precision, dev, materialize_features=(args.workload != "mvfex_pose3d"))
        P2ACT[0] = 2 if pipe.pose3d.engine().proposal_dtype() == "f16" else 4
        nb = min(B, 64)
        feat_h, bfb_h = synth.synth_features(nb, 4, seed=100 + rank)        # [nb,4,128,64,64] = 8.4 MB/frame
        if nb < B:
            feat_h, bfb_h = feat_h.repeat(B // nb, 1, 1, 1, 1), bfb_h.repeat(B // nb, 1, 1, 1, 1)
        feat_h, bfb_h = feat_h.pin_memory(), bfb_h.pin_memory()
        feat, bfb = feat_h.to(dev), bfb_h.to(dev)
        if args.workload == "pose3d":
            with torch.no_grad():
                ff = torch.relu(torch.randn_like(feat))

            def step(f=feat, b=bfb):
                preds = pipe.pose3d(f, ff, None)
                return egd.gather_rows(preds[-1].reshape(B, -1), world)
        elif args.workload == "mvfex":
            def step(f=feat, b=bfb):
                lh, lf = pipe.heatmap.forward_from_feats(f, b)
                p, _, _ = ops.get_max_preds(lh[-1].view(B * 4, 15, 64, 64), 0.5, False)
                return egd.gather_rows(p.reshape(B, -1), world)
        else:
            def step_sync(f=feat, b=bfb):
                return egd.gather_rows(pipe(f, b)["packed"], world)

            def step(f=feat, b=bfb):
                # throughput loop: independent batches alternate between two internal streams (pipeline.forward_async)
                return pipe.forward_async(f, b, world=world, lanes=args.lanes)["gathered"] if args.lanes > 1 else step_sync(f, b)
        pipe.freeze() if args.workload != "pose3d" else None
        in_bytes = feat_h.numel() * 4 + bfb_h.numel() * 4
        l2_note = "inputs %.0f MB + activations >> 126 MB L2 (no flush needed)" % (in_bytes / 1e6)

        if args.workload != "pose3d" and not args.no_parity:
            def parity_fn():
                # untimed: the step that is timed, checked against the oracle on a seeded sample of its frames
                from egorear_b200 import calib
                from oracle import parity
                pm = pipe
                out = pm(feat, bfb)
                idx = parity.sample_indices(B, 8, seed=7)
                sd_h = {k: v.cpu() for k, v in pm.heatmap.state_dict().items()}
                sd_p = {k: v.cpu() for k, v in pm.pose3d.state_dict().items()}
                return parity.hot_path_parity(out, pm.heatmap.last_anchors[0], idx, feat_h[idx], bfb_h[idx], sd_h, sd_p,
                                              calib.load_calibration(None), "ego4view_syn")
        elif not args.no_parity:
            def parity_fn():
                from egorear_b200 import calib
                from oracle import parity
                idx = parity.sample_indices(B, 8, seed=7)
                preds = pipe.pose3d(feat, ff, None)
                ii = torch.as_tensor(idx, device=dev)
                return parity.pose3d_parity(preds[-1], idx, feat[ii].cpu(), ff[ii].cpu(),
                                            {k: v.cpu() for k, v in pipe.pose3d.state_dict().items()},
                                            calib.load_calibration(None), "ego4view_syn")

        if args.workload == "mvfex_pose3d" and args.precision in ("bf16", "fp16"):
            # e2e input = what a channels-last half-precision backbone leaves on the host side of the boundary: the
            # view-major channels-last 16-bit feature maps (4.2 MB/frame instead of 8.4 MB of NCHW fp32) + the stride-32 map in 16 bits
            feat_st_h = pipe.stage_host_features(feat_h).pin_memory()
            bfb_st_h = pipe.stage_host_bottom(bfb_h).pin_memory()
            in_bytes = feat_st_h.numel() * 2 + bfb_st_h.numel() * 2
            e2e_api = ("HotPathPipeline.infer_host_batches on pinned view-major channels-last %s features + %s stride-32 map "
                       "(H2D of batch i+1 overlaps the forward of batch i)" % (args.precision, args.precision))

            def e2e_fn(n):
                return pipe.infer_host_batches(((feat_st_h, bfb_st_h) for _ in range(n)), world)
        elif args.workload == "mvfex_pose3d":
            e2e_api = "HotPathPipeline.infer_host_batches (H2D of batch i+1 overlaps the forward of batch i)"

            def e2e_fn(n):
                return pipe.infer_host_batches(((feat_h, bfb_h) for _ in range(n)), world)
        else:
            e2e_api = "module forward on freshly uploaded inputs"

            def e2e_fn(n):
                for _ in range(n):
                    f = feat_h.to(dev, non_blocking=True)
                    b = bfb_h.to(dev, non_blocking=True)
                    yield step(f, b).cpu()
    elif args.workload == "rw_e2e":
        # BASELINE config 5: images -> backbone engine -> hot path (rw cameras, per-frame device->camera transforms)
        pipe = HotPathPipeline(4, "ego4view_rw", args.precision, dev, with_backbone=True, materialize_features=False)
        P2ACT[0] = 2 if pipe.pose3d.engine().proposal_dtype() == "f16" else 4
        g = torch.Generator().manual_seed(rank)
        img_h = torch.randn(B, 4, 3, 256, 256, generator=g).pin_memory()
        ctm_h = synth.synth_coord_trans_mat(B, seed=rank).pin_memory()
        img, ctm = img_h.to(dev), ctm_h.to(dev)
        split = {}

        def step(im=img, cm=ctm):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            xh, b = pipe.backbone_staged(im)          # channels-last bf16 FPN output, handed over without re-staging
            e[1].record()
            out = egd.gather_rows(pipe.forward(None, b, cm, feat_staged=xh)["packed"], world)
            e[2].record()
            split["ev"] = e
            return out
        pipe.freeze()
        l2_note = "images %.0f MB + activations >> 126 MB L2" % ((img_h.numel() * 4 + ctm_h.numel() * 4) / 1e6)
        if not args.no_parity:
            def parity_fn():
                # hot path only (backbone parity: tests/test_gpu_backbone.py): the staged features the backbone left, as fp32 NCHW
                from egorear_b200 import calib
                from oracle import parity
                xh, b = pipe.backbone_staged(img)
                out = pipe.forward(None, b, ctm, feat_staged=xh)
                idx = parity.sample_indices(B, 8, seed=7)
                ii = torch.as_tensor(idx, device=dev)
                feat_s = xh[:, ii].permute(1, 0, 4, 2, 3).float().cpu().contiguous()
                sd_h = {k: v.cpu() for k, v in pipe.heatmap.state_dict().items() if "heatmap_estimator_" not in k}
                sd_p = {k: v.cpu() for k, v in pipe.pose3d.state_dict().items()}
                return parity.hot_path_parity(out, pipe.heatmap.last_anchors[0], idx, feat_s, b[ii].cpu(), sd_h, sd_p,
                                              calib.load_calibration(None), "ego4view_rw", ctm_h[idx])
        # end to end as the data loader sees it: DECODED uint8 frames (256x256 RGB, 0.8 MB per 4-view frame) in pinned host
        # memory -> H2D on a copy stream (overlapping the previous batch) -> resize/ToTensor/Normalize on the GPU -> backbone
        # engine -> hot path -> packed joints read back
        u8_h = torch.from_numpy(synth.synth_images(8, 256, 256, seed=rank)).repeat((B * 4 + 7) // 8, 1, 1, 1)[: B * 4]
        u8_h = u8_h.view(B, 4, 256, 256, 3).contiguous().pin_memory()
        in_bytes = u8_h.numel() + ctm_h.numel() * 4
        e2e_api = ("HotPathPipeline.infer_host_batches on pinned decoded uint8 frames [B,4,256,256,3]: H2D of batch i+1 overlaps "
                   "batch i; ops.preprocess_images + backbone engine + hot path on the device")

        def e2e_fn(n):
            yield from pipe.infer_host_batches(((u8_h, ctm_h) for _ in range(n)), world)
    elif args.workload == "generate_target":
        kp_h = torch.from_numpy(synth.synth_keypoints(B, 4, 16, seed=rank)).pin_memory()
        kp = kp_h.to(dev)
        ring = [torch.empty((B, 4, 16, 64, 64), dtype=torch.float32, device=dev) for _ in range(2)]   # 2 x 8.6 GB at B=8192
        cnt = [0]

        def step(k=kp):
            cnt[0] += 1
            return ops.generate_target_batch(k, 872, 64, 1.0, out=ring[cnt[0] & 1])
        in_bytes = kp_h.numel() * 8
        l2_note = "output ring 2 x %.1f GB >> L2" % (ring[0].numel() * 4 / 1e9)

        e2e_api = "ops.generate_target_batch on uploaded keypoints, checksum read-back"

        def e2e_fn(n):
            for _ in range(n):
                out = step(kp_h.to(dev, non_blocking=True))
                yield out[:, :, :, 0, 0].sum().cpu()                         # checksum read-back; maps stay on device
    elif args.workload == "eval_heatmap":
        from egorear_b200 import metrics
        kp = torch.from_numpy(synth.synth_keypoints(B, 4, 15, seed=rank)).to(dev)
        gt = ops.generate_target_batch(kp)                                   # [B,4,15,64,64], 2.0 GB at B=2048
        gt_h = gt.cpu().pin_memory()
        pred = gt * 0.9 + torch.randn(gt.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)) * 0.03
        in_bytes = gt_h.numel() * 4
        l2_note = "inputs 2 x %.1f GB >> L2" % (in_bytes / 1e9)

        def step(g=gt):
            return metrics._heatmap_metrics(pred, g, 1.0, "bench")[0]     # per-frame errors stay on the device
        e2e_api = "metrics.evaluate: targets uploaded from pinned host memory (they come from the data loader), per-frame errors read back"

        def e2e_fn(n):
            for _ in range(n):
                m = metrics.evaluate(pred, gt_h.to(dev, non_blocking=True), "b")
                yield torch.stack([m["b_l1_error_heatmap"], m["b_pos_l1_error_heatmap"]])
    elif args.workload == "preprocess":
        base = torch.from_numpy(synth.synth_images(8, 872, 872, seed=rank))
        img_h = base.repeat((B * 4 + 7) // 8, 1, 1, 1)[: B * 4].view(B, 4, 872, 872, 3).contiguous().pin_memory()
        img_d = img_h.to(dev)
        in_bytes = img_h.numel()
        l2_note = "decoded frames %.1f GB + output %.1f GB >> L2" % (in_bytes / 1e9, B * 4 * 3 * 256 * 256 * 4 / 1e9)

        def step(im=img_d):
            return ops.preprocess_images(im)
        e2e_api = "ops.preprocess_images on decoded uint8 frames uploaded from pinned host memory, checksum read-back"

        def e2e_fn(n):
            for _ in range(n):
                yield step(img_h.to(dev, non_blocking=True))[:, :, :, 0, 0].sum().cpu().view(1)
    elif args.workload == "eval_pose":
        from egorear_b200 import metrics
        pr, gp = synth.synth_eval_poses(B, 16, seed=rank)
        pr_h, gp_h = torch.from_numpy(pr).pin_memory(), torch.from_numpy(gp).pin_memory()
        pr_d, gp_d = pr_h.to(dev), gp_h.to(dev)
        in_bytes = gp_h.numel() * 4
        l2_note = "inputs 2 x %.0f MB + 34 MB of results > 126 MB L2" % (in_bytes / 1e6)

        def step(g=gp_d):
            return metrics._pose_metrics(pr_d, g, 10.0, 150.0, metrics._AUC_THRESHOLDS, False, "bench")[0]
        e2e_api = "metrics.evaluate_pose: ground truth uploaded from pinned host memory, the four per-sample metrics read back"

        def e2e_fn(n):
            for _ in range(n):
                m = metrics.evaluate_pose(pr_d, gp_h.to(dev, non_blocking=True), "b")
                yield torch.from_numpy(np.stack([v.astype(np.float64) for v in m.values()], 1))
    else:  # decode
        kp = torch.from_numpy(synth.synth_keypoints(B, 4, 15, seed=rank)).to(dev)
        hm = ops.generate_target_batch(kp).view(B * 4, 15, 64, 64)
        hm_h = None
        in_bytes = hm.numel() * 4
        l2_note = "input %.1f GB >> L2" % (in_bytes / 1e9)

        def step(h=hm):
            return ops.get_max_preds(h, 0.5, True)[0]

    join = pipe.join if args.workload == "mvfex_pose3d" and args.lanes > 1 else None
    step_prof = step_sync if args.workload == "mvfex_pose3d" else step
    torch.cuda.synchronize(dev)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    egd.barrier()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    sampler.start()
    time.sleep(0.05)
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    if os.environ.get("EGR_CUDA_PROFILER"):        # ncu --profile-from-start off: capture exactly the timed region
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record()
    for _ in range(args.steps):
        step()
    if join is not None:
        join()                       # the timing stream waits for every lane before the closing event
    ev1.record()
    torch.cuda.synchronize(dev)
    if os.environ.get("EGR_CUDA_PROFILER"):
        torch.cuda.cudart().cudaProfilerStop()
    egd.barrier()
    launches = _lib.launch_count() - l0
    ms = egd.max_over_ranks(ev0.elapsed_time(ev1), dev)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = world * B * args.steps / (ms / 1e3)
    # for transparency: the same loop on ONE stream (no overlap between consecutive batches)
    single = None
    if join is not None:
        n1 = max(3, min(args.steps, 20))
        for _ in range(2):           # lane 0 (the synchronous forward) has not run yet: its workspaces are allocated here, untimed
            step_sync()
        torch.cuda.synchronize(dev)
        egd.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n1):
            step_sync()
        s1.record()
        torch.cuda.synchronize(dev)
        egd.barrier()
        ms1 = egd.max_over_ranks(s0.elapsed_time(s1), dev)
        single = {"value": world * B * n1 / (ms1 / 1e3), "ms_per_step": ms1 / n1, "steps": n1}
    extra = {}
    if args.workload == "rw_e2e":
        e = split["ev"]
        extra = {"backbone_ms": e[0].elapsed_time(e[1]), "hot_path_ms": e[1].elapsed_time(e[2]),
                 "hot_path_frames_per_s": world * B / (e[1].elapsed_time(e[2]) / 1e3)}

    # ---- sustained: the same step back to back for >= sustain_s seconds (clocks settle under the ~1 kW power cap) ----
    sustained = None
    if args.sustain_s > 0 and args.workload in ("mvfex_pose3d", "mvfex", "pose3d", "rw_e2e"):
        n_s = max(args.steps, int(args.sustain_s / max(ms / 1e3 / args.steps, 1e-6)) + 1)
        samp2 = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
        torch.cuda.synchronize(dev)
        egd.barrier()
        samp2.start()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for _ in range(n_s):
            step()
        if join is not None:
            join()
        u1.record()
        torch.cuda.synchronize(dev)
        samp2.stop_flag = True
        samp2.join(timeout=2)
        egd.barrier()
        ms_s = egd.max_over_ranks(u0.elapsed_time(u1), dev)
        sustained = {"value": world * B * n_s / (ms_s / 1e3), "ms_per_step": ms_s / n_s, "steps": n_s, "seconds": ms_s / 1e3,
                     "clocks": samp2.result()}
        if args.workload in ("mvfex_pose3d", "mvfex"):
            # whole-step tensor-pipe fraction: SURVEY 8d algorithmic FLOPs (21.02 G mvfex + 1.106 G lifting per frame)
            gf = 21.02 + (1.106 if args.workload == "mvfex_pose3d" else 0.0)
            tfs = gf * 1e9 * B * n_s / (ms_s / 1e3) / 1e12
            sustained["algorithmic_tflops"] = tfs
            sustained["frac_of_sustained_tensor_peak"] = tfs / peaks["tf_sust"]

    # ---- e2e: pinned host inputs -> H2D -> hot path -> D2H of the step's result, every step (device-timed) ----
    e2e = None
    if e2e_fn is not None:
        n_e2e = max(3, min(args.steps, 10))
        for _ in e2e_fn(2):
            pass
        torch.cuda.synchronize(dev)
        egd.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in e2e_fn(n_e2e):
            pass
        e1.record()
        torch.cuda.synchronize(dev)
        egd.barrier()
        dt = egd.max_over_ranks(e0.elapsed_time(e1), dev)
        e2e = {"value": world * B * n_e2e / (dt / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(in_bytes),
               "d2h_bytes_per_step": int(r.numel() * r.element_size()), "steps": n_e2e,
               "api": e2e_api}

    # ---- per-stage timing of the same step through the library's stage profiler (rank 0) ----
    roofline, stages, stage_fracs = None, None, None
    n_prof = 5 if args.workload != "rw_e2e" else 2
    if args.workload in ("mvfex_pose3d", "mvfex", "pose3d", "rw_e2e"):
        # every rank runs the profiled steps (the step ends in the all-gather); only rank 0 records stage events
        if rank == 0:
            lib.egr_profile_enable(1)
        for _ in range(n_prof):
            step_prof()
        torch.cuda.synchronize(dev)
        egd.barrier()
    if rank == 0 and args.workload in ("mvfex_pose3d", "mvfex", "pose3d", "rw_e2e"):
        import ctypes
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(lib.egr_profile_read(buf, len(buf)))
        lib.egr_profile_enable(0)
        stages = {}
        for item in buf.value.decode().split(";"):
            if item:
                name, tot, cnt = item.split(":")
                stages[name] = float(tot) / n_prof           # ms per step
        total = sum(stages.values())
        # HotPathPipeline exports the staged copies (1); the chained workload does not materialise the NCHW features (2)
        exp = 2 if args.workload in ("mvfex_pose3d", "rw_e2e") else 1 if args.workload == "mvfex" else 0
        per_stage = {k: stage_roofline(k, v, B, act, exp, peaks) for k, v in stages.items()}
        # the roofline line is for the dominant kernel that HAS a roofline (the token phases are chains of ~15-45
        # latency-bound launches, reported in stages_ms)
        dense = [k for k in stages if per_stage[k] is not None]
        top = max(dense, key=stages.get)
        roofline = dict(per_stage[top])
        roofline.update({"kernel": top, "traffic": ncu_traffic.get(top), "traffic_source": ncu_source,
                         "peak_source": peaks["src"] + (" (burst bf16 GEMM: the stage is event-timed inside a %d-step pass)" % n_prof
                                                        if roofline["bound"] == "tensor" else " (copy)"),
                         "share_of_step": stages[top] / total, "ms_per_launch": stages[top]})
        stage_fracs = {k: {"bound": r["bound"], "frac": round(r["frac"], 3)} for k, r in per_stage.items() if r}
    elif rank == 0 and args.workload in ("generate_target", "decode"):
        t_s = ms / 1e3 / args.steps
        per_frame = 4 * 16 * 4096 * 4 + 4 * 16 * 16 if args.workload == "generate_target" else 4 * 15 * 4096 * 4 + 4 * 15 * 13
        ach = per_frame * B / t_s / 1e9
        roofline = {"kernel": args.workload, "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": ach / peaks["hbm"], "traffic": None, "peak_source": peaks["src"], "share_of_step": 1.0,
                    "ms_per_launch": t_s * 1e3}

    elif rank == 0 and args.workload == "preprocess":
        t_s = ms / 1e3 / args.steps
        per_frame = 4 * (872 * 872 * 3 + 3 * 256 * 256 * 4)        # decoded frame read once + normalised tensor written once
        ach = per_frame * B / t_s / 1e9
        roofline = {"kernel": "preprocess_kernel", "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": ach / peaks["hbm"], "traffic": None, "peak_source": peaks["src"], "share_of_step": 1.0,
                    "ms_per_launch": t_s * 1e3}
    elif rank == 0 and args.workload in ("eval_heatmap", "eval_pose"):
        t_s = ms / 1e3 / args.steps
        # algorithmic bytes: both maps of every (frame, view, joint) once + 16 B of partials | both poses + 4 doubles
        per_unit = 2 * 4 * 15 * 4096 * 4 + 4 * 15 * 16 + 8 if args.workload == "eval_heatmap" else 2 * 16 * 3 * 4 + 32
        ach = per_unit * B / t_s / 1e9
        roofline = {"kernel": "eval_heatmap_kernel (+2 reductions)" if args.workload == "eval_heatmap" else "eval_pose_kernel",
                    "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                    "traffic": None, "peak_source": peaks["src"], "share_of_step": 1.0, "ms_per_launch": t_s * 1e3}

    if rank != 0:
        egd.shutdown()
        return 0
    par = None
    if parity_fn is not None and world == 1:
        try:
            par = parity_fn()
        except Exception as e:      # the parity block must never take the measurement down with it
            par = {"error": "%s: %s" % (type(e).__name__, e)}
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.workload in ("eval_heatmap", "eval_pose"):
        cpu = cpu_baseline_eval(args.workload)
    elif world == 1 and not args.no_cpu_baseline and args.workload == "preprocess":
        cpu = cpu_baseline_preprocess()
    elif world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision] if args.workload in ("mvfex_pose3d", "mvfex", "pose3d", "rw_e2e") else "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "frames_per_gpu_per_step": B, "precision": args.precision + ("" if args.workload not in ("mvfex_pose3d", "pose3d") else
                                                       " (pose3d proposal branch: %s)" % pipe.pose3d.engine().proposal_dtype()),
                       "weights": "random-init (name-seeded), shipped architecture", "l2": l2_note,
                       "streams": ("%d lanes: independent batches alternate between internal streams" % args.lanes)
                       if join is not None else "1",
                       "collective": "1 NCCL all-gather of the packed joints per step" if world > 1 else "none (1 GPU)",
                       "options": dict(kv.split("=", 1) for kv in args.opt) if args.opt else "library defaults"},
            "clocks": sampler.result(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "parity": par, "sustained": sustained, "stages_ms": stages, "stage_roofline": stage_fracs,
            "single_stream": single}
    line.update(extra)
    emit(line)
    egd.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
