"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY.

Compile the REFERENCE's own query kernels for sm_100a, from the source where it lies under
/root/reference, into oracle/_ref/libref_query_K<K>.so (git-ignored, travels to the GPU box).

The reference keeps its CUDA as a Python string handed to pycuda.compiler.SourceModule
(models/neural_points/query_point_indices_worldcoords.py:136-683).  This recipe extracts that string
at build time into a temporary directory, substitutes `KN` (= opt.K, :138), appends
oracle/ref_launcher.inc (our launchers) and runs nvcc with nvcc's defaults for everything PyCUDA
leaves at default (-fmad=true, IEEE division).  No reference source is copied into the repository:
only the compiled .so is kept.  Where /root/reference is absent (the GPU box) the prebuilt .so is used.
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("SGN_REFERENCE_ROOT", "/root/reference")
REF_FILE = os.path.join(REF_ROOT, "models", "neural_points", "query_point_indices_worldcoords.py")
# the perspective-frustum querier (--wcoord_query 0): the same recipe over query_point_indices.py:130-600 and ref_launcher_pers.inc
REF_FILE_PERS = os.path.join(REF_ROOT, "models", "neural_points", "query_point_indices.py")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def so_path(K=8, pers=False):
    return os.path.join(OUT, f"libref_query{'_pers' if pers else ''}_K{K}.so")


def extract_source(K, pers=False):
    text = open(REF_FILE_PERS if pers else REF_FILE, encoding="utf-8").read()
    start = text.index("SourceModule(")
    end = text.index('""", no_extern_c=True)', start)
    body = text[start:end]
    first = body.index('"""') + 3
    body = body[first:]
    # the `""" + str(self.opt.K)\n + """` splice that defines KN
    body = re.sub(r'"""\s*\+\s*str\(self\.opt\.K\)\s*\+\s*"""', str(K), body)
    assert '"""' not in body, "unexpected extra string splice in the reference source"
    return body


def build(K=8, force=False, keep_sass=False, pers=False):
    """Returns the .so path, or None when the reference tree is not present and nothing was prebuilt."""
    target = so_path(K, pers)
    ref_file = REF_FILE_PERS if pers else REF_FILE
    if not os.path.exists(ref_file):
        return target if os.path.exists(target) else None
    launcher = os.path.join(HERE, "ref_launcher_pers.inc" if pers else "ref_launcher.inc")
    if (not force and os.path.exists(target)
            and os.path.getmtime(target) >= max(os.path.getmtime(ref_file), os.path.getmtime(launcher))):
        return target
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        cu = os.path.join(tmp, "ref_query.cu")
        with open(cu, "w", encoding="utf-8") as f:
            f.write(extract_source(K, pers))
            f.write("\n")
            f.write(open(launcher).read())
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
               "-w", "-o", target, cu]
        subprocess.check_call(cmd)
        if keep_sass:
            sass = subprocess.check_output(["cuobjdump", "-sass", target]).decode()
            open(os.path.join(OUT, f"ref_query{'_pers' if pers else ''}_K{K}.sass"), "w").write(sass)
    return target


if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    print(build(k, force=True, keep_sass=True))
    print(build(k, force=True, keep_sass=True, pers=True))
