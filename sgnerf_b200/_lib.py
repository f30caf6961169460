"""ctypes binding of libsgnerf_b200.so (include/sgnerf_b200.h).

There is no fallback: if the shared library is missing, loading raises; if a call fails, the
library's error text is raised as RuntimeError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libsgnerf_b200.so")

c_void = C.c_void_p
c_i64 = C.c_int64
c_int = C.c_int
c_f32 = C.c_float
c_u64 = C.c_uint64
c_size = C.c_size_t


class SgnGridCfg(C.Structure):
    _fields_ = [("origin", c_f32 * 3), ("vsize", c_f32 * 3), ("dim", C.c_int32 * 3), ("query_size", C.c_int32 * 3),
                ("P", C.c_int32), ("max_o", C.c_int32), ("seconds_claim", c_u64), ("seconds_fill", c_u64)]


class SgnPersCfg(C.Structure):
    _fields_ = [("shift", c_f32 * 3), ("vsize", c_f32 * 3), ("dim", C.c_int32 * 3), ("vscale", C.c_int32 * 3), ("kernel_size", C.c_int32 * 3),
                ("query_size", C.c_int32 * 3), ("ray_vsize", c_f32 * 3), ("P", C.c_int32), ("SR", C.c_int32), ("K", C.c_int32), ("NN", C.c_int32),
                ("inverse", C.c_int32), ("radius2", c_f32), ("depth2", c_f32), ("seconds_insert", c_u64), ("seconds_query", c_u64)]


class SgnAggCfg(C.Structure):
    _fields_ = [("feat_dim", C.c_int32), ("num_feat_freqs", C.c_int32), ("dist_xyz_freq", C.c_int32),
                ("num_viewdir_freqs", C.c_int32), ("width", C.c_int32), ("n_block1", C.c_int32),
                ("n_block2_bpnet", C.c_int32), ("label_dim", C.c_int32), ("n_block3", C.c_int32),
                ("n_color", C.c_int32), ("act_super", C.c_int32), ("leaky_slope", c_f32)]


class SgnPointTables(C.Structure):
    _fields_ = [("xyz", c_void), ("embedding", c_void), ("color", c_void), ("dir", c_void), ("conf", c_void),
                ("label_emb", c_void), ("N", c_i64)]


class SgnPointGrads(C.Structure):
    _fields_ = [("embedding", c_void), ("color", c_void), ("dir", c_void), ("conf", c_void)]


# name -> (restype, argtypes); mirrors include/sgnerf_b200.h one to one
SIGNATURES = {
    "sgn_last_error": (C.c_char_p, []),
    "sgn_version": (c_int, []),
    "sgn_launch_count": (c_u64, []),
    "sgn_grid_workspace_bytes": (c_int, [c_i64, C.POINTER(SgnGridCfg), C.POINTER(c_size), C.POINTER(c_size)]),
    "sgn_grid_build": (c_int, [c_void, c_i64, c_i64, C.POINTER(SgnGridCfg), c_void, c_size, c_void, c_size,
                               C.POINTER(c_void), c_void]),
    "sgn_grid_build_flags": (c_int, [c_void, c_i64, c_i64, C.POINTER(SgnGridCfg), c_void, c_size, c_void, c_size, c_int,
                                     C.POINTER(c_void), c_void]),
    "sgn_grid_destroy": (c_int, [c_void]),
    "sgn_grid_buffer": (c_int, [c_void, c_int, C.POINTER(c_void), C.POINTER(c_i64)]),
    "sgn_query": (c_int, [c_void, c_void, c_void, c_void, c_int, c_i64, c_int, c_int, c_int, c_int, c_f32,
                          c_void, c_void, c_void, c_u64, c_void, c_void, c_void, c_void, c_void, c_void]),
    "sgn_query_frame": (c_int, [c_void, c_void, c_void, c_void, c_int, c_i64, c_int, c_int, c_int, c_int, c_f32,
                                c_void, c_void, c_void, c_u64, c_void, c_void, c_void, c_void, c_void, c_void]),
    "sgn_query_march_mode": (c_int, [c_int]),
    "sgn_gather_rows": (c_int, [c_void, c_int, c_void, c_i64, c_void, c_void]),
    "sgn_agg_num_layers": (c_int, [C.POINTER(SgnAggCfg)]),
    "sgn_agg_layer_shape": (c_int, [C.POINTER(SgnAggCfg), c_int, C.POINTER(c_int), C.POINTER(c_int)]),
    "sgn_agg_workspace_bytes": (c_int, [C.POINTER(SgnAggCfg), c_i64, c_i64, c_int, c_int, c_int, c_int, C.POINTER(c_size)]),
    "sgn_agg_forward": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointTables),
                                c_void, c_void, c_void, c_void, c_void, c_i64, c_int, c_int, c_int, c_int,
                                c_void, c_void, c_void, c_void, c_void, c_void, c_size, c_void]),
    "sgn_agg_forward_cached": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointTables),
                                       c_void, c_void, c_void, c_void, c_void, c_i64, c_int, c_int, c_int, c_int,
                                       c_void, c_void, c_void, c_void, c_void, c_void, c_size, c_void, c_void]),
    "sgn_agg_forward_frame": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointTables),
                                      c_void, c_void, c_void, c_void, c_void, c_i64, c_int, c_int, c_int, c_int,
                                      c_void, c_void, c_void, c_void, c_void, c_void, c_void, c_size, c_void, c_void]),
    "sgn_agg_forward_frame_masked": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointTables),
                                             c_void, c_void, c_void, c_void, c_void, c_void, c_i64, c_int, c_int, c_int, c_int,
                                             c_void, c_void, c_void, c_void, c_void, c_void, c_void, c_size, c_void, c_void]),
    "sgn_agg_point_cache_bytes": (c_int, [C.POINTER(SgnAggCfg), c_i64, C.POINTER(c_size)]),
    "sgn_agg_point_cache_build": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(c_void), C.POINTER(SgnPointTables), c_void, c_size, c_void]),
    "sgn_agg_point_cache_update": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(SgnPointTables), c_void, c_size, c_void, c_i64, c_void]),
    "sgn_agg_backward": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointTables),
                                 c_void, c_void, c_void, c_void, c_void, c_i64, c_int, c_int, c_void, c_void,
                                 C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointGrads), c_void, c_size, c_void]),
    "sgn_agg_backward_prec": (c_int, [C.POINTER(SgnAggCfg), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointTables),
                                      c_void, c_void, c_void, c_void, c_void, c_i64, c_int, c_int, c_int, c_void, c_void,
                                      C.POINTER(c_void), C.POINTER(c_void), C.POINTER(SgnPointGrads), c_void, c_size, c_void]),
    "sgn_agg_kernel_timing": (c_int, [c_int]),
    "sgn_agg_kernel_timing_read": (c_int, [C.POINTER(c_f32), C.POINTER(c_int)]),
    "sgn_ray_dist": (c_int, [c_void, c_void, c_f32, c_int, c_i64, c_int, c_void, c_void]),
    "sgn_composite_forward": (c_int, [c_void, c_void, c_void, c_void, c_int, c_i64, c_int, c_void, c_void, c_void,
                                      c_void, c_void, c_void]),
    "sgn_composite_backward": (c_int, [c_void, c_void, c_void, c_void, c_int, c_i64, c_int, c_void, c_void, c_void,
                                       c_void, c_void, c_void]),
    "sgn_render_composite": (c_int, [c_void, c_void, c_void, c_void, c_f32, c_int, c_void, c_int, c_i64, c_int, c_void, c_void, c_void, c_void,
                                     c_void]),
    "sgn_render_composite_depth": (c_int, [c_void, c_void, c_void, c_void, c_f32, c_int, c_void, c_int, c_i64, c_int, c_void, c_void, c_void,
                                           c_void, c_void]),
    "sgn_probe_outputs": (c_int, [c_void, c_void, c_void, c_void, c_void, c_void, C.POINTER(SgnPointTables), c_int, c_i64, c_int, c_int,
                                  c_void, c_void, c_void, c_void, c_void, c_void, c_void, c_void]),
    "sgn_fill_invalid": (c_int, [c_void, c_void, c_i64, c_int, c_void, c_void, c_void, c_void]),
    "sgn_pers_query_bytes": (c_int, [c_i64, c_i64, C.POINTER(SgnPersCfg), C.POINTER(c_size)]),
    "sgn_pers_query": (c_int, [c_void, c_i64, c_void, c_i64, C.POINTER(SgnPersCfg), c_void, c_size, c_void, c_void, c_void, c_void]),
    "sgn_voxel_downsample_bytes": (c_int, [c_i64, C.POINTER(c_size)]),
    "sgn_voxel_downsample": (c_int, [c_void, c_i64, C.POINTER(c_f32), C.POINTER(c_f32), c_int, c_void, c_size, c_void, c_void, c_void, c_void, c_void]),
    "sgn_query_vox_grid": (c_int, [c_void, c_i64, c_void, c_int, C.POINTER(c_f32), c_f32, c_void, c_void]),
    "sgn_loss_hit_count": (c_int, [c_void, c_i64, c_void, c_void]),
    "sgn_loss_forward_backward": (c_int, [c_void, c_void, c_void, c_void, c_i64, c_int, c_int, c_void, c_f32, c_f32, c_f32, c_f32, c_void, c_void,
                                          c_void, c_void]),
    "sgn_adam_step_count": (c_int, [c_void, c_void]),
    "sgn_adam_rows_multi": (c_int, [c_int, C.POINTER(c_void), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(C.c_int32), c_void,
                                    c_i64, c_f32, c_f32, c_f32, c_f32, c_void, c_f32, c_int, c_void]),
    "sgn_rows_union_bytes": (c_int, [c_i64, C.POINTER(c_size)]),
    "sgn_rows_union": (c_int, [c_void, c_i64, c_void, c_void, c_void, c_size, c_void]),
    "sgn_rows_pack": (c_int, [c_int, C.POINTER(c_void), C.POINTER(C.c_int32), c_void, c_void, c_i64, c_void, c_int, c_int, c_void]),
    "sgn_adam_dense_multi": (c_int, [c_int, C.POINTER(c_void), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(c_i64), c_f32, c_f32, c_f32,
                                     c_f32, c_void, c_f32, c_int, c_void]),
    "sgn_adam_mark_rows": (c_int, [c_void, c_i64, c_void, c_void]),
    "sgn_adam_rows_list": (c_int, [c_int, C.POINTER(c_void), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(c_void), C.POINTER(C.c_int32), c_void, c_void,
                                   c_void, c_void, c_i64, c_f32, c_f32, c_f32, c_f32, c_void, c_f32, c_int, c_void]),
    "sgn_adam_rows": (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_int, c_f32, c_f32, c_f32, c_f32, c_void, c_f32, c_int, c_void]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: build it with `python -m sgnerf_b200.build` "
                               "(sgnerf_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().sgn_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args), name)
