"""ctypes front-end of oracle/libgs_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may
import this module.  The product package (gsorb_slam_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgs_oracle.so")
_lib = None

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gs_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "cpu"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.gso_forward.restype = C.c_void_p
        L.gso_forward.argtypes = [C.c_int, C.c_int, C.c_int, _f, C.c_int, C.c_int, _f, _f, _f, _f, _f,
                                  C.c_float, _f, _f, _f, _f, _f, C.c_float, C.c_float, _f, _f, _i, _i]
        L.gso_free.argtypes = [C.c_void_p]
        L.gso_num_rendered.restype = C.c_longlong
        L.gso_num_rendered.argtypes = [C.c_void_p]
        L.gso_get_geometry.argtypes = [C.c_void_p, _f, _f, _f, _f, _f, _u]
        L.gso_get_binning.argtypes = [C.c_void_p, _u, C.POINTER(C.c_uint64), _u]
        L.gso_get_image_state.argtypes = [C.c_void_p, _f, _u]
        L.gso_backward.argtypes = [C.c_void_p] + [_f] * 5 + [C.c_float] + [_f] * 5 + [C.c_float, C.c_float] + [_f] * 10
        L.gso_visible_filter.argtypes = [C.c_int, C.c_int, C.c_int, _f, _f, C.c_float, _f, _f, _f,
                                         C.c_float, C.c_float, _i]
        L.gso_mark_visible.argtypes = [C.c_int, _f, _f, _f, C.POINTER(C.c_uint8)]
        L.gso_knn.argtypes = [C.c_int, _f, _f]
        L.gso_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _fp(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class OracleFrame:
    """One forward pass of the CPU restatement, keeping its state for backward / inspection."""

    def __init__(self, *, width, height, means3D, opacities, background, viewmatrix, projmatrix,
                 tanfovx, tanfovy, colors=None, shs=None, sh_degree=0, scales=None, rotations=None,
                 cov3D=None, scale_modifier=1.0, campos=None):
        L = lib()
        self.W, self.H = int(width), int(height)
        self.P = int(means3D.shape[0])
        self.M = 0 if shs is None else int(shs.shape[1])
        self.D = int(sh_degree)
        self.means3D, self.opacities = _c(means3D), _c(opacities).reshape(-1)
        self.colors, self.shs, self.scales = _c(colors), _c(shs), _c(scales)
        self.rotations, self.cov3D = _c(rotations), _c(cov3D)
        self.background = _c(background)
        self.view, self.proj = _c(viewmatrix).reshape(16), _c(projmatrix).reshape(16)
        self.campos = _c(campos if campos is not None else np.zeros(3))
        self.tanfovx, self.tanfovy, self.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
        self.color = np.zeros((3, self.H, self.W), np.float32)
        self.depth = np.zeros((1, self.H, self.W), np.float32)
        self.radii = np.zeros(self.P, np.int32)
        nr = C.c_int(0)
        self._h = L.gso_forward(self.P, self.D, self.M, _fp(self.background), self.W, self.H,
                                _fp(self.means3D), _fp(self.shs), _fp(self.colors), _fp(self.opacities),
                                _fp(self.scales), self.scale_modifier, _fp(self.rotations), _fp(self.cov3D),
                                _fp(self.view), _fp(self.proj), _fp(self.campos), self.tanfovx, self.tanfovy,
                                _fp(self.color), _fp(self.depth), self.radii.ctypes.data_as(_i), C.byref(nr))
        self.num_rendered = int(L.gso_num_rendered(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().gso_free(self._h)
            self._h = None

    def geometry(self):
        P = self.P
        out = dict(depths=np.zeros(P, np.float32), means2D=np.zeros((P, 2), np.float32),
                   cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
                   rgb=np.zeros((P, 3), np.float32), tiles_touched=np.zeros(P, np.uint32))
        lib().gso_get_geometry(self._h, _fp(out["depths"]), _fp(out["means2D"]), _fp(out["cov3D"]),
                               _fp(out["conic_opacity"]), _fp(out["rgb"]),
                               out["tiles_touched"].ctypes.data_as(_u))
        return out

    def binning(self):
        R = self.num_rendered
        tiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        out = dict(point_list=np.zeros(R, np.uint32), keys=np.zeros(R, np.uint64),
                   ranges=np.zeros((tiles, 2), np.uint32))
        lib().gso_get_binning(self._h, out["point_list"].ctypes.data_as(_u),
                              out["keys"].ctypes.data_as(C.POINTER(C.c_uint64)),
                              out["ranges"].ctypes.data_as(_u))
        return out

    def image_state(self):
        out = dict(final_T=np.zeros((self.H, self.W), np.float32), n_contrib=np.zeros((self.H, self.W), np.uint32))
        lib().gso_get_image_state(self._h, _fp(out["final_T"]), out["n_contrib"].ctypes.data_as(_u))
        return out

    def backward(self, dL_dpix):
        P, M = self.P, self.M
        dL = _c(dL_dpix)
        g = dict(dL_dmean2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
                 dL_dopacity=np.zeros(P, np.float32), dL_dcolor=np.zeros((P, 3), np.float32),
                 dL_dmean3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
                 dL_dsh=np.zeros((P, max(M, 1), 3), np.float32) if self.shs is not None else None,
                 dL_dscale=np.zeros((P, 3), np.float32) if self.scales is not None else None,
                 dL_drot=np.zeros((P, 4), np.float32) if self.rotations is not None else None)
        lib().gso_backward(self._h, _fp(self.background), _fp(self.means3D), _fp(self.shs), _fp(self.colors),
                           _fp(self.scales), self.scale_modifier, _fp(self.rotations), _fp(self.cov3D),
                           _fp(self.view), _fp(self.proj), _fp(self.campos), self.tanfovx, self.tanfovy,
                           _fp(dL), _fp(g["dL_dmean2D"]), _fp(g["dL_dconic"]), _fp(g["dL_dopacity"]),
                           _fp(g["dL_dcolor"]), _fp(g["dL_dmean3D"]), _fp(g["dL_dcov3D"]), _fp(g["dL_dsh"]),
                           _fp(g["dL_dscale"]), _fp(g["dL_drot"]))
        return g


def frame_from_scene(scene, **overrides) -> OracleFrame:
    cam = scene.cam
    kw = dict(width=cam.width, height=cam.height, means3D=scene.means3D, opacities=scene.opacities,
              background=scene.background, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
              tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, colors=scene.colors, scales=scene.scales,
              rotations=scene.rotations, campos=cam.campos)
    kw.update(overrides)
    return OracleFrame(**kw)


def visible_filter(*, width, height, means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy,
                   scale_modifier=1.0):
    P = means3D.shape[0]
    radii = np.zeros(P, np.int32)
    lib().gso_visible_filter(P, int(width), int(height), _fp(_c(means3D)), _fp(_c(scales)), float(scale_modifier),
                             _fp(_c(rotations)), _fp(_c(viewmatrix).reshape(16)), _fp(_c(projmatrix).reshape(16)),
                             float(tanfovx), float(tanfovy), radii.ctypes.data_as(_i))
    return radii


def mark_visible(means3D, viewmatrix, projmatrix):
    P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    lib().gso_mark_visible(P, _fp(_c(means3D)), _fp(_c(viewmatrix).reshape(16)), _fp(_c(projmatrix).reshape(16)),
                           out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def knn_mean_dist2(points):
    pts = _c(points)
    out = np.zeros(pts.shape[0], np.float32)
    lib().gso_knn(pts.shape[0], _fp(pts), _fp(out))
    return out


def num_threads() -> int:
    return int(lib().gso_num_threads())
