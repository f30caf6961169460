"""Build libsgnerf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m sgnerf_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libsgnerf_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "sgnerf_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, hdr_time):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    sp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(sp), hdr_time):
        return obj, False
    subprocess.check_call([NVCC] + FLAGS + ["-c", sp, "-o", obj])
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = _newest_header()
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_time), sources()))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(SO):
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs)
        if verbose:
            print("linked", SO)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
