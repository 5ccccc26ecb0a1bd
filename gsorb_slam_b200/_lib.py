"""ctypes binding of libgsb.so (include/gsb.h) -- the only way the Python host side reaches
the kernels.  There is NO fallback: if the library is missing the import of any operator
raises, loudly, with the build command."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSB_LIB: developer override (e.g. the tuning build libgsb_tune.so of `make -C gsorb_slam_b200/csrc tune`)
LIB_PATH = os.environ.get("GSB_LIB") or os.path.join(_HERE, "libgsb.so")

GSB_OK = 0
ERRORS = {-1: "GSB_ERR_INVALID_ARGUMENT", -2: "GSB_ERR_CUDA", -3: "GSB_ERR_WORKSPACE", -4: "GSB_ERR_OVERFLOW",
          -5: "GSB_ERR_UNSUPPORTED"}

_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class RasterArgs(C.Structure):
    """gsb_raster_args"""
    _fields_ = [("P", _i), ("D", _i), ("M", _i), ("width", _i), ("height", _i), ("background", _vp),
                ("means3D", _vp), ("shs", _vp), ("colors_precomp", _vp), ("opacities", _vp), ("scales", _vp),
                ("scale_modifier", _f), ("rotations", _vp), ("cov3D_precomp", _vp), ("viewmatrix", _vp),
                ("projmatrix", _vp), ("cam_pos", _vp), ("tan_fovx", _f), ("tan_fovy", _f), ("prefiltered", _i),
                ("tile_row_begin", _i), ("tile_row_end", _i)]


class GradOutputs(C.Structure):
    """gsb_grad_outputs"""
    _fields_ = [(n, _vp) for n in ("dL_dmean2D", "dL_dconic", "dL_dopacity", "dL_dcolor", "dL_dmean3D", "dL_dcov3D",
                                   "dL_dsh", "dL_dscale", "dL_drot")]


class MapUpdate(C.Structure):
    """gsb_map_update (groups: means, rgb, logit opacities, log scales, unnormalised quaternions)"""
    _fields_ = [("Tcw", _vp), ("params", _vp * 5), ("exp_avg", _vp * 5), ("exp_avg_sq", _vp * 5), ("grads", _vp * 5),
                ("lr", C.c_float * 5), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("step", _ll),
                ("dL_dTcw", _vp), ("max_scalar", C.c_float), ("w_scalar", C.c_float), ("w_long", C.c_float), ("reg_terms", _vp)]


# name -> (restype, argtypes); every symbol include/gsb.h declares
SIGNATURES = {
    "gsb_version": (_i, []),
    "gsb_last_error": (C.c_char_p, []),
    "gsb_geometry_bytes": (_sz, [_i]),
    "gsb_image_bytes": (_sz, [_i, _i]),
    "gsb_binning_bytes": (_sz, [_ll]),
    "gsb_workspace_query": (_i, [_i, _i, _i, _ll, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "gsb_forward": (_i, [C.POINTER(RasterArgs), ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp, _vp, _vp, _vp, _vp]),
    "gsb_forward_ws": (_i, [C.POINTER(RasterArgs), _vp, _sz, _vp, _sz, _ll, _vp, _sz, _vp, _vp, _vp, _vp]),
    "gsb_num_rendered": (_ll, [_vp, _vp]),
    "gsb_backward": (_i, [C.POINTER(RasterArgs), _ll, _vp, _vp, _vp, _vp, _vp, C.POINTER(GradOutputs), _vp]),
    "gsb_forward_fused_ws": (_i, [C.POINTER(RasterArgs), _vp, _sz, _vp, _sz, _ll, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "gsb_backward_fused": (_i, [C.POINTER(RasterArgs), _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(GradOutputs), _vp, _i, _vp]),
    "gsb_backward_fused_update": (_i, [C.POINTER(RasterArgs), _vp, _vp, _vp, _vp, _vp, _vp, _i, C.POINTER(MapUpdate), _vp]),
    "gsb_backward_fused_pose": (_i, [C.POINTER(RasterArgs), _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "gsb_visible_filter": (_i, [C.POINTER(RasterArgs), _vp, _vp]),
    "gsb_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "gsb_knn_workspace_bytes": (_sz, [_i]),
    "gsb_knn_mean_dist2": (_i, [_i, _vp, _vp, _vp, _sz, _vp]),
    "gsb_prologue": (_i, [_i] + [_vp] * 10),
    "gsb_prologue_backward": (_i, [_i] + [_vp] * 15),
    "gsb_pose_grad": (_i, [_i, _vp, _vp, _vp, _vp]),
    "gsb_adam_step": (_i, [_ll, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, _ll, _vp]),
    "gsb_scale_regulariser": (_i, [_i, _vp, _f, _f, _f, _vp, _vp, _vp]),
    "gsb_loss_scratch_bytes": (_sz, [_i, _i]),
    "gsb_tracking_loss": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "gsb_mapping_loss": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gsb_backproject_scratch_bytes": (_sz, [_i, _i]),
    "gsb_backproject": (_i, [_i, _i, _vp, _vp, _vp, _f, _f, _f, _f, C.POINTER(_f), _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gsb_prune_scratch_bytes": (_sz, [_i]),
    "gsb_low_opacity_keep": (_i, [_i, _vp, _f, _vp, _vp]),
    "gsb_prune_rows": (_i, [_i, _vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), _vp, _vp, _sz, _vp]),
    "gsb_exchange_sync_bytes": (_sz, [_i]),
    "gsb_exchange_allreduce": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), _ll, _i, _i, _vp]),
    "gsb_adam_step_groups": (_i, [_i, C.POINTER(_ll), C.POINTER(_f), _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double, _ll, _vp]),
    "gsb_host_scratch_bytes": (_sz, [_i, _i, _i, _i, _ll]),
    "gsb_forward_backward_host": (_ll, [C.POINTER(RasterArgs), _ll, _vp, _vp, _vp, _vp, C.POINTER(GradOutputs), _vp, _sz, _vp]),
    "gsb_forward_backward_host_async": (_i, [C.POINTER(RasterArgs), _ll, _vp, _vp, _vp, _vp, C.POINTER(GradOutputs), _vp, _sz, _vp, _vp]),
    "gsb_debug_image_state": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "gsb_debug_blended_pairs": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "gsb_debug_binning_state": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "gsb_debug_geometry_state": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "gsb_launch_count_reset": (_ll, []),
    "gsb_num_stages": (_i, []),
    "gsb_stage_name": (C.c_char_p, [_i]),
    "gsb_profile_begin": (_i, []),
    "gsb_profile_end": (_i, [C.POINTER(_f), C.POINTER(_i)]),
}

_lib = None


class GsbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def lib():
    """Load libgsb.so (once).  Raises if it has not been built: there is no CPU / torch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing -- build it with `make -C gsorb_slam_b200/csrc -j8` "
                              "(or python -c 'import __graft_entry__ as g; g.build()'); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the ABI and the header drifted apart
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> int:
    """Negative status -> the exception type the reference surface would raise
    (std::invalid_argument / AT_ERROR -> ValueError; everything else RuntimeError)."""
    if rc < 0:
        msg = lib().gsb_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        raise GsbError(int(rc), msg)
    return rc
