"""Recipe for oracle/_ref: the UNMODIFIED reference files of the hot path, copied from where they lie.

    python oracle/make_ref.py            # needs /root/reference (the build container); a no-op elsewhere

The reference is pure Python on this path (no C/C++ to compile), so "building" oracle/_ref means copying the
handful of source files the path imports into oracle/_ref/ (git-ignored: reference sources never enter the
history; NOT gpurun-ignored: the copy travels to the GPU box like a built .so).  Nothing is edited.  Used by
    * tests/   -- the reference's own loop code driven over the drop-in modules, signature / state_dict parity
    * bench.py -- the `--impl reference` arm and the cpu_baseline leg time these modules (kind "reference")
and by nothing on the product path (oracle/ is test infrastructure, see oracle/diffreg_oracle.py).
"""
import os
import shutil
import sys

REF = os.environ.get("DRG_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXP_2D3D = "Diff-Reg-2d3d/experiments/2d3dmatr.rgbdv2.stage4.level3.stage1"
FILES = [
    # 4DMatch tree: the matching / procrustes / pipeline modules and what `import models.pipeline` pulls in
    "Diff-Reg-4dmatch/models/__init__.py",
    "Diff-Reg-4dmatch/models/matching.py",
    "Diff-Reg-4dmatch/models/procrustes.py",
    "Diff-Reg-4dmatch/models/position_encoding.py",
    "Diff-Reg-4dmatch/models/pipeline.py",
    "Diff-Reg-4dmatch/models/transformer.py",
    "Diff-Reg-4dmatch/models/blocks.py",
    "Diff-Reg-4dmatch/models/backbone.py",
    "Diff-Reg-4dmatch/kernels/kernel_points.py",
    "Diff-Reg-4dmatch/lib/__init__.py",
    "Diff-Reg-4dmatch/lib/ply.py",
    # 3DMatch tree (its pipeline.py needs open3d: not importable)
    "Diff-Reg-3dmatch/models/__init__.py",
    "Diff-Reg-3dmatch/models/matching.py",
    "Diff-Reg-3dmatch/models/procrustes.py",
    "Diff-Reg-3dmatch/models/position_encoding.py",
    # 2D-3D tree
    EXP_2D3D + "/matching.py",
    EXP_2D3D + "/procrustes.py",
    EXP_2D3D + "/position_encoding.py",
    "Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py",
    # the 2D-3D fusion module and the pure-torch vision3d layer files it imports
    EXP_2D3D + "/fusion_module.py",
    "Diff-Reg-2d3d/vision3d/layers/transformer.py",
    "Diff-Reg-2d3d/vision3d/layers/embedding.py",
] + ["Diff-Reg-2d3d/vision3d/layers/basic_layers/" + f for f in
     ("__init__.py", "builder.py", "depthwise_conv.py", "monte_carlo_dropout.py", "norm.py", "separable_conv.py", "utils.py")]


def make_ref(verbose=True):
    """Copy FILES from the reference checkout into oracle/_ref/.  Returns True if oracle/_ref is complete."""
    if not os.path.isdir(REF):
        ok = all(os.path.exists(os.path.join(OUT, f)) for f in FILES)
        if verbose:
            print(f"oracle/make_ref: {REF} not present; oracle/_ref {'already complete' if ok else 'absent'}")
        return ok
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(src):
            if f.endswith("__init__.py"):
                open(dst, "a").close()
                continue
            raise FileNotFoundError(src)
        shutil.copyfile(src, dst)
    if verbose:
        print(f"oracle/make_ref: {len(FILES)} reference files -> {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
