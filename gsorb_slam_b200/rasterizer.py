"""Host-side mirror of GSORB-SLAM's rasterizer operator surface, over the C ABI of libgsb.so.

Mirrors (same names, argument meaning, defaults and error behaviour):

* ``GaussianRasterizationSettings``  -- include/Rasterizer.cuh:79-91
* ``GaussianRasterizer.forward / Visable / mark_visible`` -- include/Rasterizer.cuh:284-380
* ``rasterize_gaussians`` / ``_RasterizeGaussians`` -- src/Rasterizer.cu:8-73, include/Rasterizer.cuh:127-282
* ``distCUDA2`` -- src/spatial.cu:15-27

torch is used for device memory, streams and autograd plumbing only; every computation goes
through ``gsb_*`` entry points (ctypes).  There is no eager / CPU fallback: tensors that are
not on a CUDA device raise, and a missing libgsb.so raises at first use.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import GradOutputs, RasterArgs


@dataclass
class GaussianRasterizationSettings:
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    camera_center: torch.Tensor
    prefiltered: bool = False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None   # the reference passes the null data_ptr of an empty tensor (Rasterizer.cuh:320-334)
    return t.data_ptr()


def _f32(t: Optional[torch.Tensor], what: str) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: libgsb has no CPU path")
    return t.detach().to(torch.float32).contiguous()


def _defined(t) -> bool:
    return t is not None and t.numel() > 0


def _make_args(means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, rs: GaussianRasterizationSettings,
               keep: list) -> RasterArgs:
    """Marshal tensors into gsb_raster_args; `keep` collects the contiguous copies so they outlive the call."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise ValueError("means3D must have dimensions (num_points, 3)")   # AT_ERROR, src/Rasterizer.cu:158-160
    t = dict(means3D=_f32(means3D, "means3D"), shs=_f32(sh, "shs"), colors_precomp=_f32(colors_precomp, "colors_precomp"),
             opacities=_f32(opacities, "opacities"), scales=_f32(scales, "scales"), rotations=_f32(rotations, "rotations"),
             cov3D_precomp=_f32(cov3D, "cov3D_precomp"), background=_f32(rs.bg, "bg"),
             viewmatrix=_f32(rs.viewmatrix, "viewmatrix"), projmatrix=_f32(rs.projmatrix, "projmatrix"),
             cam_pos=_f32(rs.camera_center, "camera_center"))
    keep.extend(v for v in t.values() if v is not None)
    a = RasterArgs()
    a.P = int(means3D.size(0))
    a.D = int(rs.sh_degree)
    a.M = int(sh.size(1)) if _defined(sh) else 0          # src/Rasterizer.cu:183-187
    a.width, a.height = int(rs.image_width), int(rs.image_height)
    for k, v in t.items():
        setattr(a, k, _ptr(v))
    a.scale_modifier = float(rs.scale_modifier)
    a.tan_fovx, a.tan_fovy = float(rs.tanfovx), float(rs.tanfovy)
    a.prefiltered = int(bool(rs.prefiltered))
    return a


class _Blob:
    """A caller-owned scratch blob handed to gsb_forward through the allocator callback
    (the role of resizeFunctional, src/Rasterizer.cu:127-134 -- but not zero-filled)."""

    def __init__(self, device):
        self.t = torch.empty(0, dtype=torch.uint8, device=device)
        self.device = device

        def cb(_user, nbytes):
            self.t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=self.device)
            return self.t.data_ptr()
        self.cb = _lib.ALLOC_FN(cb)


class _RasterizeGaussians(torch.autograd.Function):
    """include/Rasterizer.cuh:127-282.  forward -> (color [3,H,W], radii [P] i32, depth [1,H,W]);
    backward consumes only d(color) (Rasterizer.cuh:210): radii and median depth carry no gradient."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
        L = _lib.lib()
        rs = raster_settings
        dev = means3D.device
        if not means3D.is_cuda:
            raise RuntimeError("means3D must be a CUDA tensor: libgsb has no CPU path")
        keep: list = []
        with torch.cuda.device(dev):
            a = _make_args(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, keep)
            H, W, P = a.height, a.width, a.P
            color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
            depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            geom, binning, img = _Blob(dev), _Blob(dev), _Blob(dev)
            R = _lib.check(L.gsb_forward(C.byref(a), geom.cb, None, binning.cb, None, img.cb, None, color.data_ptr(),
                                         depth.data_ptr(), _ptr(radii), _stream()))
        ctx.raster_settings = rs
        ctx.num_rendered = R
        ctx.blobs = (geom.t, binning.t, img.t)
        ctx.has = (_defined(sh), _defined(colors_precomp), _defined(scales), _defined(rotations), _defined(cov3Ds_precomp))
        ctx.save_for_backward(means3D, sh if _defined(sh) else None, colors_precomp if _defined(colors_precomp) else None,
                              opacities, scales if _defined(scales) else None, rotations if _defined(rotations) else None,
                              cov3Ds_precomp if _defined(cov3Ds_precomp) else None, radii)
        ctx.mark_non_differentiable(radii, depth)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, _grad_depth):
        L = _lib.lib()
        means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, radii = ctx.saved_tensors
        rs = ctx.raster_settings
        dev = means3D.device
        P = means3D.size(0)
        M = sh.size(1) if sh is not None else 0
        keep: list = []
        with torch.cuda.device(dev):
            a = _make_args(means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, rs, keep)
            dpix = grad_out_color.detach().to(torch.float32).contiguous()
            e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
            g_means3D, g_means2D, g_conic, g_opac = e(P, 3), e(P, 3), e(P, 2, 2), e(P, 1)
            g_color, g_cov3D = e(P, 3), e(P, 6)
            g_sh = e(P, M, 3) if M else None
            g_scale = e(P, 3) if scales is not None else None
            g_rot = e(P, 4) if rotations is not None else None
            g = GradOutputs(dL_dmean2D=g_means2D.data_ptr(), dL_dconic=g_conic.data_ptr(), dL_dopacity=g_opac.data_ptr(),
                            dL_dcolor=g_color.data_ptr(), dL_dmean3D=g_means3D.data_ptr(), dL_dcov3D=g_cov3D.data_ptr(),
                            dL_dsh=_ptr(g_sh), dL_dscale=_ptr(g_scale), dL_drot=_ptr(g_rot))
            geom, binning, img = ctx.blobs
            _lib.check(L.gsb_backward(C.byref(a), ctx.num_rendered, radii.data_ptr(), geom.data_ptr(), binning.data_ptr(),
                                      img.data_ptr(), dpix.data_ptr(), C.byref(g), _stream()))
        has_sh, has_col, has_scale, has_rot, has_cov = ctx.has
        # order of the forward arguments (Rasterizer.cuh:259-266)
        return (g_means3D, g_means2D, g_sh if has_sh else None, g_color if has_col else None,
                g_opac.reshape(opacities.shape), g_scale if has_scale else None, g_rot if has_rot else None,
                g_cov3D if has_cov else None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, device_num,
                        raster_settings):
    """src/Rasterizer.cu:8-73 (device_num kept for signature parity; the device is taken from means3D)."""
    del device_num
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


class GaussianRasterizer(torch.nn.Module):
    """include/Rasterizer.cuh:284-380."""

    def __init__(self, raster_settings: Optional[GaussianRasterizationSettings] = None):
        super().__init__()
        self.raster_settings_ = raster_settings

    @property
    def raster_settings(self):
        return self.raster_settings_

    def mark_visible(self, positions: torch.Tensor) -> torch.Tensor:
        L = _lib.lib()
        rs = self.raster_settings_
        with torch.no_grad(), torch.cuda.device(positions.device):
            p = _f32(positions, "positions")
            v, pm = _f32(rs.viewmatrix, "viewmatrix"), _f32(rs.projmatrix, "projmatrix")
            P = int(positions.size(0))
            present = torch.empty((P,), dtype=torch.bool, device=positions.device)
            _lib.check(L.gsb_mark_visible(P, _ptr(p), _ptr(v), _ptr(pm), _ptr(present), _stream()))
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, device_num: int = 0):
        if _defined(shs) == _defined(colors_precomp):
            raise ValueError("Please provide exactly one of either SHs or precomputed colors!")
        if ((_defined(scales) or _defined(rotations)) and _defined(cov3D_precomp)) or \
                (not _defined(scales) and not _defined(rotations) and not _defined(cov3D_precomp)):
            raise ValueError("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        color, radii, depth = rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                                  cov3D_precomp, device_num, self.raster_settings_)
        return color, radii, depth

    def Visable(self, means3D, opacities=None, scales=None, rotations=None, device_num: int = 0):
        """Radii-only projection (Rasterizer.cuh:351-376 -> filter_radii, src/Rasterizer.cu:75-122)."""
        del opacities, device_num
        L = _lib.lib()
        rs = self.raster_settings_
        keep: list = []
        with torch.no_grad(), torch.cuda.device(means3D.device):
            a = _make_args(means3D, None, None, None, scales, rotations, None, rs, keep)
            radii = torch.empty((a.P,), dtype=torch.int32, device=means3D.device)
            _lib.check(L.gsb_visible_filter(C.byref(a), _ptr(radii), _stream()))
        return (radii,)

    visible = Visable


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """src/spatial.cu:15-27: mean squared distance of every point to its 3 nearest neighbours."""
    L = _lib.lib()
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor: libgsb has no CPU path")
    with torch.no_grad(), torch.cuda.device(points.device):
        p = points.detach().to(torch.float32).contiguous()
        P = int(p.size(0))
        means = torch.zeros((P,), dtype=torch.float32, device=p.device)
        if P:
            nbytes = int(L.gsb_knn_workspace_bytes(P))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
            _lib.check(L.gsb_knn_mean_dist2(P, p.data_ptr(), means.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return means
