"""Thin, autograd-free driver of the C ABI: one object = one rasterized frame.

Used by the parity tests, bench.py and __graft_entry__.smoke(); it is the Python twin of
what src/Rasterizer.cu:136-297 does with libtorch (allocate outputs, hand scratch blobs to
the library, keep them for backward).  Two modes:

* ``sync_free=False``: gsb_forward with allocator callbacks, one stream sync to learn
  num_rendered (the reference's contract);
* ``sync_free=True``: gsb_forward_ws over caller-sized workspaces (capacity ``max_rendered``),
  nothing blocks; ``num_rendered`` is read (and overflow detected) on demand.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GradOutputs, RasterArgs


def _dev(a, device, dtype=torch.float32):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype).contiguous()


def _p(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


class Frame:
    def __init__(self, *, width, height, means3D, opacities, background, viewmatrix, projmatrix, tanfovx, tanfovy,
                 colors=None, shs=None, sh_degree=0, scales=None, rotations=None, cov3D=None, scale_modifier=1.0,
                 campos=None, device="cuda:0", sync_free=False, max_rendered=None, run=True, fused=False, tile_rows=None):
        self.L = _lib.lib()
        self.device = torch.device(device)
        d = self.device
        self.W, self.H = int(width), int(height)
        self.means3D = _dev(means3D, d)
        self.P = int(self.means3D.shape[0])
        self.opacities = _dev(opacities, d).reshape(-1)
        self.colors, self.shs, self.scales = _dev(colors, d), _dev(shs, d), _dev(scales, d)
        self.rotations, self.cov3D = _dev(rotations, d), _dev(cov3D, d)
        self.M = 0 if self.shs is None else int(self.shs.shape[1])
        self.D = int(sh_degree)
        self.bg = _dev(background, d)
        self.view, self.proj = _dev(viewmatrix, d).reshape(16), _dev(projmatrix, d).reshape(16)
        self.campos = _dev(campos if campos is not None else np.zeros(3, np.float32), d)
        self.tanfovx, self.tanfovy, self.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
        self.tile_rows = tile_rows  # (begin, end) tile rows rendered by this frame (tile-row shard); None = whole image
        self.fused = bool(fused)   # five-channel pass: RGB + [z_cam, 1] of the depth pass (gsb_forward_fused_ws)
        self.sync_free = bool(sync_free) or self.fused
        self.max_rendered = int(max_rendered) if max_rendered is not None else 4 * self.P + 1024
        self.num_rendered = None
        self._args = self._make_args()
        # a band render leaves the other rows untouched: zero-filled so that bands can be summed
        mk = torch.zeros if tile_rows is not None else torch.empty
        self.color = mk((3, self.H, self.W), dtype=torch.float32, device=d)
        self.depth = mk((1, self.H, self.W), dtype=torch.float32, device=d)
        self.radii = torch.empty((self.P,), dtype=torch.int32, device=d)
        self.depth_sil = mk((2, self.H, self.W), dtype=torch.float32, device=d) if self.fused else None
        self.geom = self.binning = self.img = None
        self._grads = None
        if run:
            self.forward()

    def _make_args(self) -> RasterArgs:
        a = RasterArgs()
        a.P, a.D, a.M, a.width, a.height = self.P, self.D, self.M, self.W, self.H
        a.background, a.means3D, a.shs, a.colors_precomp = _p(self.bg), _p(self.means3D), _p(self.shs), _p(self.colors)
        a.opacities, a.scales, a.scale_modifier = _p(self.opacities), _p(self.scales), self.scale_modifier
        a.rotations, a.cov3D_precomp = _p(self.rotations), _p(self.cov3D)
        a.viewmatrix, a.projmatrix, a.cam_pos = _p(self.view), _p(self.proj), _p(self.campos)
        a.tan_fovx, a.tan_fovy, a.prefiltered = self.tanfovx, self.tanfovy, 0
        if self.tile_rows is not None:
            # an empty band must not be (0, 0): that pair is the C ABI's "whole image" (include/gsb.h, gsb_raster_args)
            from .distributed import band_for_rasterizer
            a.tile_row_begin, a.tile_row_end = band_for_rasterizer(self.tile_rows, (self.H + 15) // 16)
        return a

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _alloc_ws(self):
        L, d = self.L, self.device
        if self.geom is None:
            self.geom = torch.empty(int(L.gsb_geometry_bytes(self.P)), dtype=torch.uint8, device=d)
            self.img = torch.empty(int(L.gsb_image_bytes(self.W, self.H)), dtype=torch.uint8, device=d)
            self.binning = torch.empty(int(L.gsb_binning_bytes(self.max_rendered)), dtype=torch.uint8, device=d)

    def forward(self):
        L = self.L
        with torch.cuda.device(self.device):
            if self.fused:
                self._alloc_ws()
                _lib.check(L.gsb_forward_fused_ws(C.byref(self._args), self.geom.data_ptr(), self.geom.numel(),
                                                  self.binning.data_ptr(), self.binning.numel(), self.max_rendered,
                                                  self.img.data_ptr(), self.img.numel(), self.color.data_ptr(),
                                                  self.depth_sil.data_ptr(), self.depth.data_ptr(), _p(self.radii), self._stream()))
                self.num_rendered = None
            elif self.sync_free:
                self._alloc_ws()
                _lib.check(L.gsb_forward_ws(C.byref(self._args), self.geom.data_ptr(), self.geom.numel(),
                                            self.binning.data_ptr(), self.binning.numel(), self.max_rendered,
                                            self.img.data_ptr(), self.img.numel(), self.color.data_ptr(),
                                            self.depth.data_ptr(), _p(self.radii), self._stream()))
                self.num_rendered = None
            else:
                blobs = {}

                def mk(name):
                    def cb(_u, n):
                        blobs[name] = torch.empty(max(int(n), 1), dtype=torch.uint8, device=self.device)
                        return blobs[name].data_ptr()
                    return _lib.ALLOC_FN(cb)
                g, b, i = mk("g"), mk("b"), mk("i")
                R = _lib.check(L.gsb_forward(C.byref(self._args), g, None, b, None, i, None, self.color.data_ptr(),
                                             self.depth.data_ptr(), _p(self.radii), self._stream()))
                self.geom, self.binning, self.img = blobs["g"], blobs["b"], blobs["i"]
                self.num_rendered = int(R)
        return self

    def rendered(self) -> int:
        """num_rendered (synchronises in sync-free mode; raises GSB_ERR_OVERFLOW if the capacity was exceeded)."""
        if self.num_rendered is None:
            with torch.cuda.device(self.device):
                self.num_rendered = int(_lib.check(self.L.gsb_num_rendered(self.geom.data_ptr(), self._stream())))
        return self.num_rendered

    def alloc_grads(self):
        P, M, d = self.P, self.M, self.device
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=d)
        g = dict(dL_dmean2D=e(P, 3), dL_dconic=e(P, 4), dL_dopacity=e(P), dL_dcolor=e(P, 3), dL_dmean3D=e(P, 3),
                 dL_dcov3D=e(P, 6), dL_dsh=e(P, M, 3) if M else None,
                 dL_dscale=e(P, 3) if self.scales is not None else None,
                 dL_drot=e(P, 4) if self.rotations is not None else None)
        go = GradOutputs(**{k: _p(v) for k, v in g.items()})
        self._grads = (g, go)
        return g

    def backward_fused(self, dL_dcolor, dL_ddepth_sil, reuse_outputs=False, z_attached=False):
        """Backward of the five-channel pass: the SUM of the RGB pass' and the depth pass' gradients, plus
        ``dL_dzcolor`` [P], the gradient of the depth pass' z_cam colour."""
        if self._grads is None or not reuse_outputs:
            self.alloc_grads()
        g, go = self._grads
        dC, dD = _dev(dL_dcolor, self.device), _dev(dL_ddepth_sil, self.device)
        g["dL_dzcolor"] = torch.empty(self.P, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.gsb_backward_fused(C.byref(self._args), _p(self.radii), self.geom.data_ptr(), self.binning.data_ptr(),
                                                 self.img.data_ptr(), dC.data_ptr(), dD.data_ptr(), C.byref(go),
                                                 g["dL_dzcolor"].data_ptr(), 1 if z_attached else 0, self._stream()))
        return g

    def backward(self, dL_dpix, reuse_outputs=False):
        if self._grads is None or not reuse_outputs:
            self.alloc_grads()
        g, go = self._grads
        dL = _dev(dL_dpix, self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.gsb_backward(C.byref(self._args), -1 if self.num_rendered is None else self.num_rendered,
                                           _p(self.radii), self.geom.data_ptr(), self.binning.data_ptr(),
                                           self.img.data_ptr(), dL.data_ptr(), C.byref(go), self._stream()))
        return g

    # ---- introspection (gsb_debug_*) ----
    def image_state(self):
        d = self.device
        tiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        fT = torch.empty((self.H, self.W), dtype=torch.float32, device=d)
        nc = torch.empty((self.H, self.W), dtype=torch.int32, device=d)
        rg = torch.empty((tiles, 2), dtype=torch.int32, device=d)
        with torch.cuda.device(d):
            _lib.check(self.L.gsb_debug_image_state(self.img.data_ptr(), self.W, self.H, fT.data_ptr(), nc.data_ptr(),
                                                    rg.data_ptr(), self._stream()))
        return dict(final_T=fT, n_contrib=nc, ranges=rg)

    def blended_pairs(self) -> int:
        """(pixel, splat) pairs the forward pass blended (set bits of its hit words): the blend kernels' unit of work."""
        c = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.gsb_debug_blended_pairs(self.geom.data_ptr(), self.binning.data_ptr(), self.img.data_ptr(), self.W, self.H,
                                                      c.data_ptr(), self._stream()))
        return int(c.item())

    def binning_state(self):
        R = self.rendered()
        pl = torch.empty((R,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.gsb_debug_binning_state(self.geom.data_ptr(), self.binning.data_ptr(), R, _p(pl), self._stream()))
        return dict(point_list=pl)

    def geometry_state(self):
        P, d = self.P, self.device
        out = dict(depths=torch.empty(P, dtype=torch.float32, device=d), means2D=torch.empty((P, 2), dtype=torch.float32, device=d),
                   conic_opacity=torch.empty((P, 4), dtype=torch.float32, device=d),
                   tiles_touched=torch.empty(P, dtype=torch.int32, device=d))
        with torch.cuda.device(d):
            _lib.check(self.L.gsb_debug_geometry_state(self.geom.data_ptr(), P, _p(out["depths"]), _p(out["means2D"]),
                                                       _p(out["conic_opacity"]), _p(out["tiles_touched"]), self._stream()))
        return out


def frame_from_scene(scene, **overrides) -> Frame:
    cam = scene.cam
    kw = dict(width=cam.width, height=cam.height, means3D=scene.means3D, opacities=scene.opacities,
              background=scene.background, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
              tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, colors=scene.colors, scales=scene.scales,
              rotations=scene.rotations, campos=cam.campos)
    kw.update(overrides)
    return Frame(**kw)
