"""ORACLE recipe (test infrastructure): leave a runnable file copy of the reference's Python path under oracle/_ref/.

The reference is pure Python, so there is nothing to compile; what the GPU box lacks is the checkout itself
(/root/reference exists only in the build container).  This script copies the files the hot path imports —
pose_estimation/**/*.py, the camera calibration JSONs, configs/*.yaml and generate_heatmap.py — from the checkout,
where they lie, into oracle/_ref/, which is git-ignored (it never enters the history) but not gpurun-ignored, so it
travels with the snapshot like a built .so.  bench.py's CPU arm (`--impl reference` and the `cpu_baseline` leg) then
times the reference's own modules on the GPU box's host cores (`kind: "reference"`), through oracle/ref_import.py's
shims for the un-vendored dependencies (mmcv, timm, natsort).  Run by __graft_entry__.build() when the checkout is present.

    python oracle/make_ref.py [/path/to/EgoRear]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP_EXT = (".py", ".json", ".yaml", ".yml")


def make_ref(src="/root/reference", dst=os.path.join(HERE, "_ref")):
    if not os.path.isdir(os.path.join(src, "pose_estimation")):
        return None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for sub in ("pose_estimation", "configs"):
        for root, _, files in os.walk(os.path.join(src, sub)):
            for f in files:
                if f.endswith(KEEP_EXT):
                    rel = os.path.relpath(os.path.join(root, f), src)
                    os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(src, rel), os.path.join(dst, rel))
                    n += 1
    shutil.copyfile(os.path.join(src, "generate_heatmap.py"), os.path.join(dst, "generate_heatmap.py"))
    with open(os.path.join(dst, "PROVENANCE.txt"), "w") as fh:
        fh.write("file copy of %s made by oracle/make_ref.py (%d files); git-ignored, test infrastructure only\n" % (src, n + 1))
    return dst


if __name__ == "__main__":
    out = make_ref(*(sys.argv[1:2]))
    print("oracle/_ref:", out)
