"""Stage the reference files the GPU box needs into baseline/_ref/ (git-ignored, travels with the gpurun snapshot).

    python baseline/stage_reference.py          # build container only: needs /root/reference

The reference is not a package (no setup.py / pyproject), so `pip install --target baseline/_ref /root/reference` has nothing to
install; this copies, unmodified, the source files of the path under test and of its real caller:

    models/neural_points_volumetric_model.py   NeuralPointsRayMarching.forward (:435-671), fill_invalid (:158-195)
    models/base_rendering_model.py, base_model.py   loss terms (:543-641), found_funcs
    models/aggregators/point_aggregators.py    the reference PointAggregator (CPU baseline / torch-on-GPU baseline / golden vectors)
    models/rendering/diff_ray_marching.py, diff_render_func.py     ray_march, ray generation, render / blend functions
    models/helpers/networks.py, geometrics.py  positional_encoding, init_seq
    utils/format.py, spherical.py, util.py

Nothing under baseline/_ref/ is product code: it is read by tests/ref_import.py (the drop-in test under the reference's own caller),
bench.py --impl reference and bench.py's reference_gpu leg.  Reference sources are never copied into tracked paths.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SGN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "models/__init__.py", "models/base_model.py", "models/base_rendering_model.py", "models/neural_points_volumetric_model.py",
    "models/aggregators/__init__.py", "models/aggregators/point_aggregators.py",
    "models/rendering/__init__.py", "models/rendering/diff_ray_marching.py", "models/rendering/diff_render_func.py",
    "models/helpers/__init__.py", "models/helpers/networks.py", "models/helpers/geometrics.py",
    "utils/format.py", "utils/spherical.py", "utils/util.py",
]


def stage(verbose=False):
    """Returns baseline/_ref if it holds the staged files (copying them first when the reference tree is present), else None."""
    if os.path.isdir(REF):
        for rel in FILES:
            src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                shutil.copy2(src, dst)
                if verbose:
                    print("staged", rel)
    ok = all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)
    return DST if ok else None


if __name__ == "__main__":
    out = stage(verbose=True)
    print("baseline/_ref:", out if out else "not available (no reference tree here)")
    sys.exit(0 if out else 1)
