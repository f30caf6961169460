"""CPU: the C-ABI library loads and exports every symbol include/sgnerf_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sgnerf_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["sgn_grid_build", "sgn_query", "sgn_agg_forward", "sgn_agg_backward", "sgn_composite_forward",
              "sgn_composite_backward", "sgn_ray_dist", "sgn_fill_invalid", "sgn_gather_rows", "sgn_last_error"]:
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib]).decode()
    exported = set(re.findall(r" T (sgn_[a-z0-9_]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_binding_matches_header(built_lib):
    from sgnerf_b200 import _lib
    lib = _lib.load()
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert lib.sgn_version() >= 100
    assert isinstance(lib.sgn_last_error(), bytes)


def test_library_is_sm100a_only(built_lib):
    out = subprocess.check_output(["cuobjdump", "-lelf", built_lib]).decode()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_validation_without_gpu(built_lib):
    """Pure host-side checks run before any CUDA call."""
    from sgnerf_b200 import _lib, ops
    lib = _lib.load()
    cfg = ops.agg_cfg()
    shapes = ops.agg_layer_shapes(cfg)
    assert shapes == [(284, 256), (256, 256), (263, 256), (256, 256), (256, 1), (280, 128), (128, 128), (128, 128), (128, 3)]
    sem = ops.agg_cfg(n_block2_bpnet=1, label_dim=96)
    assert ops.agg_layer_shapes(sem)[2] == (352, 256)
    assert sum(i * o + o for i, o in ops.agg_layer_shapes(sem)) == 432132
    assert sum(i * o + o for i, o in shapes) == 341764
    bad = ops.agg_cfg(width=100)
    assert lib.sgn_agg_num_layers(ctypes.byref(bad)) < 0
    assert b"width" in lib.sgn_last_error()
    g = _lib.SgnGridCfg()
    assert lib.sgn_grid_workspace_bytes(10, ctypes.byref(g), None, None) < 0


def test_no_cpu_fallback_in_product():
    """The product package must never import the oracle."""
    pkg = os.path.join(ROOT, "sgnerf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_cpu_tensors_are_rejected(built_lib):
    import torch
    from sgnerf_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.composite(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.ones(2, 4, dtype=torch.bool))


def test_header_is_plain_c(tmp_path):
    """include/sgnerf_b200.h is the C-ABI contract: it must compile as C99 on its own (no C++ or CUDA types in the signatures) and a C
    program must link against the shared library with nothing but the header."""
    src = tmp_path / "use.c"
    src.write_text('#include "sgnerf_b200.h"\nint main(void) { return sgn_version() >= 100 && sgn_last_error() != 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "use.o")])
