"""ctypes front-end of oracle/_ref/libgsref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

libgsref.so is the UNMODIFIED reference CUDA rasterizer (+ simple_knn) compiled in place from
/root/reference by oracle/Makefile, behind the extern "C" shim oracle/ref_shim.cu.  This
module drives it the way src/Rasterizer.cu:136-297 does (zero-filled growable scratch
tensors, zero-initialised gradient buffers) with torch owning the device memory.

Used by: tests/ (-m gpu parity against the real reference), tests/golden/make_golden.py
(fixture generation) and bench.py --impl reference.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgsref.so")
_lib = None
_ALLOC = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)
_vp = C.c_void_p


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_forward.restype = C.c_int
        L.ref_forward.argtypes = [_ALLOC, _vp, _ALLOC, _vp, _ALLOC, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int,
                                  C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp, _vp,
                                  C.c_float, C.c_float, C.c_int, _vp, _vp, _vp]
        L.ref_backward.restype = None
        L.ref_backward.argtypes = [C.c_int] * 4 + [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp,
                                                   _vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, _vp] + [_vp] * 10
        L.ref_visible_filter.restype = None
        L.ref_visible_filter.argtypes = [_ALLOC, _vp, _ALLOC, _vp, _ALLOC, _vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                         _vp, _vp, C.c_float, _vp, _vp, _vp, C.c_float, C.c_float, C.c_int, _vp]
        L.ref_mark_visible.restype = None
        L.ref_mark_visible.argtypes = [C.c_int, _vp, _vp, _vp, _vp]
        L.ref_knn.restype = None
        L.ref_knn.argtypes = [C.c_int, _vp, _vp]
        sz = C.POINTER(C.c_size_t)
        L.ref_image_state_offsets.argtypes = [C.c_size_t, sz, sz, sz]
        L.ref_binning_state_offsets.argtypes = [C.c_size_t, sz, sz, sz, sz]
        L.ref_geometry_state_offsets.argtypes = [C.c_size_t] + [sz] * 8
        L.ref_sync.restype = C.c_int
        _lib = L
    return _lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _dev(a, dtype=torch.float32):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype).contiguous()


class _Scratch:
    """resizeFunctional of src/Rasterizer.cu:127-134: a byte tensor that is zero-filled on growth.  ``keep``: the tensor
    survives the call and is handed back as is when it is large enough (the "kernels only" timing variant of SURVEY.md 8d:
    no allocation and no fill inside the timed region)."""

    def __init__(self, keep: bool = False):
        self.t = torch.empty(0, dtype=torch.uint8, device="cuda")

        def cb(_user, nbytes):
            if not (keep and self.t.numel() >= int(nbytes)):
                self.t = torch.zeros(int(nbytes), dtype=torch.uint8, device="cuda")
            return self.t.data_ptr()
        self.cb = _ALLOC(cb)


class RefFrame:
    """Forward (+ optional backward) through the reference kernels; tensors stay on the GPU."""

    def __init__(self, *, width, height, means3D, opacities, background, viewmatrix, projmatrix, tanfovx,
                 tanfovy, colors=None, shs=None, sh_degree=0, scales=None, rotations=None, cov3D=None,
                 scale_modifier=1.0, campos=None, run=True, reuse=False):
        self.reuse = bool(reuse)   # keep outputs / scratch / gradient tensors across calls ("kernels only" timing)
        self._grads = None
        self.W, self.H = int(width), int(height)
        self.means3D = _dev(means3D)
        self.P = int(self.means3D.shape[0])
        self.opacities = _dev(opacities).reshape(-1)
        self.colors, self.shs, self.scales = _dev(colors), _dev(shs), _dev(scales)
        self.rotations, self.cov3D = _dev(rotations), _dev(cov3D)
        self.M = 0 if self.shs is None else int(self.shs.shape[1])
        self.D = int(sh_degree)
        self.bg = _dev(background)
        self.view, self.proj = _dev(viewmatrix).reshape(16), _dev(projmatrix).reshape(16)
        self.campos = _dev(campos if campos is not None else np.zeros(3, np.float32))
        self.tanfovx, self.tanfovy, self.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
        self.num_rendered = 0
        if run:
            self.forward()

    def forward(self):
        """RasterizeGaussiansCUDA (src/Rasterizer.cu:136-217): fresh zeroed outputs + scratch per call."""
        if not (self.reuse and hasattr(self, "color")):
            self.color = torch.zeros((3, self.H, self.W), dtype=torch.float32, device="cuda")
            self.depth = torch.zeros((1, self.H, self.W), dtype=torch.float32, device="cuda")
            self.radii = torch.zeros(self.P, dtype=torch.int32, device="cuda")
            self.geom, self.binning, self.img = _Scratch(self.reuse), _Scratch(self.reuse), _Scratch(self.reuse)
        self.num_rendered = lib().ref_forward(
            self.geom.cb, None, self.binning.cb, None, self.img.cb, None, self.P, self.D, self.M, _p(self.bg),
            self.W, self.H, _p(self.means3D), _p(self.shs), _p(self.colors), _p(self.opacities), _p(self.scales),
            self.scale_modifier, _p(self.rotations), _p(self.cov3D), _p(self.view), _p(self.proj), _p(self.campos),
            self.tanfovx, self.tanfovy, 0, _p(self.color), _p(self.depth), _p(self.radii))
        return self.num_rendered

    def backward(self, dL_dpix):
        """RasterizeGaussiansBackwardCUDA (src/Rasterizer.cu:220-297): zero-initialised gradients."""
        P, M = self.P, self.M
        dL = _dev(dL_dpix)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device="cuda")
        if self.reuse and self._grads is not None:
            g = self._grads
            for v in g.values():   # the kernels accumulate with atomics: the zero state is part of their contract
                v.zero_()
        else:
            g = dict(dL_dmean2D=z(P, 3), dL_dconic=z(P, 4), dL_dopacity=z(P), dL_dcolor=z(P, 3), dL_dmean3D=z(P, 3),
                     dL_dcov3D=z(P, 6), dL_dsh=z(P, max(M, 1), 3), dL_dscale=z(P, 3), dL_drot=z(P, 4))
            self._grads = g if self.reuse else None
        lib().ref_backward(P, self.D, self.M, self.num_rendered, _p(self.bg), self.W, self.H, _p(self.means3D),
                           _p(self.shs), _p(self.colors), _p(self.scales), self.scale_modifier, _p(self.rotations),
                           _p(self.cov3D), _p(self.view), _p(self.proj), _p(self.campos), self.tanfovx, self.tanfovy,
                           _p(self.radii), _p(self.geom.t), _p(self.binning.t), _p(self.img.t), _p(dL),
                           _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolor"]),
                           _p(g["dL_dmean3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscale"]),
                           _p(g["dL_drot"]))
        return g

    # ---- slices of the reference's opaque state blobs (layout: ref_shim.cu offsets) ----
    def image_state(self):
        n = self.W * self.H
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib().ref_image_state_offsets(n, C.byref(a), C.byref(b), C.byref(c))
        raw = self.img.t
        tiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        return dict(final_T=raw[a.value:a.value + 4 * n].view(torch.float32).reshape(self.H, self.W),
                    n_contrib=raw[b.value:b.value + 4 * n].view(torch.int32).reshape(self.H, self.W),
                    ranges=raw[c.value:c.value + 8 * tiles].view(torch.int32).reshape(tiles, 2))

    def binning_state(self):
        R = self.num_rendered
        o = [C.c_size_t() for _ in range(4)]
        lib().ref_binning_state_offsets(R, *[C.byref(x) for x in o])
        raw = self.binning.t
        return dict(point_list=raw[o[0].value:o[0].value + 4 * R].view(torch.int32),
                    keys=raw[o[2].value:o[2].value + 8 * R].view(torch.int64))

    def geometry_state(self):
        P = self.P
        o = [C.c_size_t() for _ in range(8)]
        lib().ref_geometry_state_offsets(P, *[C.byref(x) for x in o])
        raw = self.geom.t
        f = lambda k, n: raw[o[k].value:o[k].value + 4 * n].view(torch.float32)
        return dict(depths=f(0, P), means2D=f(3, 2 * P).reshape(P, 2), cov3D=f(4, 6 * P).reshape(P, 6),
                    conic_opacity=f(5, 4 * P).reshape(P, 4), rgb=f(6, 3 * P).reshape(P, 3),
                    tiles_touched=raw[o[7].value:o[7].value + 4 * P].view(torch.int32))


def frame_from_scene(scene, **overrides) -> RefFrame:
    cam = scene.cam
    kw = dict(width=cam.width, height=cam.height, means3D=scene.means3D, opacities=scene.opacities,
              background=scene.background, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
              tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, colors=scene.colors, scales=scene.scales,
              rotations=scene.rotations, campos=cam.campos)
    kw.update(overrides)
    return RefFrame(**kw)


def visible_filter(*, width, height, means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy,
                   scale_modifier=1.0):
    m, s, r = _dev(means3D), _dev(scales), _dev(rotations)
    v, p = _dev(viewmatrix).reshape(16), _dev(projmatrix).reshape(16)
    P = m.shape[0]
    radii = torch.zeros(P, dtype=torch.int32, device="cuda")
    a, b, c = _Scratch(), _Scratch(), _Scratch()
    lib().ref_visible_filter(a.cb, None, b.cb, None, c.cb, None, P, 0, int(width), int(height), _p(m), _p(s),
                             float(scale_modifier), _p(r), _p(v), _p(p), float(tanfovx), float(tanfovy), 0, _p(radii))
    torch.cuda.synchronize()
    return radii


def mark_visible(means3D, viewmatrix, projmatrix):
    m, v, p = _dev(means3D), _dev(viewmatrix).reshape(16), _dev(projmatrix).reshape(16)
    out = torch.zeros(m.shape[0], dtype=torch.uint8, device="cuda")
    lib().ref_mark_visible(m.shape[0], _p(m), _p(v), _p(p), _p(out))
    torch.cuda.synchronize()
    return out


def knn_mean_dist2(points):
    pts = _dev(points)
    out = torch.zeros(pts.shape[0], dtype=torch.float32, device="cuda")
    lib().ref_knn(pts.shape[0], _p(pts), _p(out))
    torch.cuda.synchronize()
    return out
