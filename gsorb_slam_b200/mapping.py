"""Fused render-optimise step over the C ABI: the hot loop of ``Render::RenderForFrame``
(src/Render.cc:402-493) with the Gaussian parameters, their Adam state and the gradient block
resident on the GPU in packed ``[14, P]`` form.

One ``step()`` = prologue (pose transform + sigmoid / normalize / exp, src/Render.cc:750-759)
-> rasterizer forward -> caller-supplied dL/dpixel -> rasterizer backward -> prologue backward
(+ camera-pose gradient dL/dTcw) -> all-reduce of the gradient block over the
ranks (distributed.py) -> Adam (``torch::optim::Adam`` semantics of src/Gaussian.cc:131-175:
betas (0.9, 0.999), eps 1e-15, per-group learning rates of Examples/RGB-D/replica.yaml:95-99).
With world size 1 the step is the reference's single-view step.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import GradOutputs, RasterArgs
from .distributed import BLOCK_ROWS, GROUPS, GradBlock, SymmetricExchange, allreduce_gradients, world

# Examples/RGB-D/replica.yaml:95-99 (Mapping.lrs*)
DEFAULT_LR = {"means": 1e-4, "rgb": 2.5e-3, "quats": 1e-3, "opacity": 0.05, "scales": 1e-3}


class MapOptimizer:
    def __init__(self, means, rgb, logit_opacities, log_scales, unnorm_quats, *, width, height, tanfovx, tanfovy,
                 projmatrix, background=None, lr: Optional[Dict[str, float]] = None, betas=(0.9, 0.999), eps=1e-15,
                 device="cuda:0", max_rendered: Optional[int] = None):
        self.L = _lib.lib()
        self.dev = torch.device(device)
        d = self.dev
        t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(d, torch.float32)
        self.P = int(t(means).shape[0])
        P = self.P
        self.params = GradBlock(P, d)   # same packed layout as the gradient block
        self.params["means"].copy_(t(means)); self.params["rgb"].copy_(t(rgb))
        self.params["opacity"].copy_(t(logit_opacities).reshape(P, 1)); self.params["scales"].copy_(t(log_scales))
        self.params["quats"].copy_(t(unnorm_quats))
        # world size > 1: the gradient block lives in a symmetric allocation and is all-reduced by ONE libgsb kernel
        # over NVLink (distributed.SymmetricExchange); NCCL only if peer mapping is unavailable
        self.exchange = None
        if world()[1] > 1 and self.dev.type == "cuda":
            try:
                self.exchange = SymmetricExchange(BLOCK_ROWS * P, d)
            except Exception:
                self.exchange = None
        self.grads = GradBlock(P, d, storage=self.exchange.alloc(BLOCK_ROWS * P) if self.exchange else None)
        self.exp_avg, self.exp_avg_sq = GradBlock(P, d), GradBlock(P, d)
        self.lr = dict(DEFAULT_LR if lr is None else lr)
        self.betas, self.eps, self.t = betas, float(eps), 0
        self.W, self.H = int(width), int(height)
        self.tanfovx, self.tanfovy = float(tanfovx), float(tanfovy)
        self.proj = t(projmatrix).reshape(16).contiguous()
        # default mode of Render::StartSplatting: identity view, means pre-transformed (src/Render.cc:748-754)
        self.view = torch.eye(4, device=d).reshape(16).contiguous()
        self.campos = torch.zeros(3, device=d)
        self.bg = t(background if background is not None else np.zeros(3, np.float32))
        # activated temporaries + their gradients
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=d)
        self.means_cam, self.opac, self.rot, self.scales = e(P, 3), e(P), e(P, 4), e(P, 3)
        self.g_means_cam, self.g_opac, self.g_rot, self.g_scales = e(P, 3), e(P), e(P, 4), e(P, 3)
        self.g_side = e(P, 13)   # mean2D 3 | conic 4 | cov3D 6: reference outputs nobody consumes here
        self.dTcw = e(3, 4)
        self.color, self.depth = e(3, self.H, self.W), e(1, self.H, self.W)
        self.radii = torch.empty(P, dtype=torch.int32, device=d)
        self.max_rendered = int(max_rendered) if max_rendered else 4 * P + 4096
        L = self.L
        u8 = lambda n: torch.empty(int(n), dtype=torch.uint8, device=d)
        self.geom, self.img = u8(L.gsb_geometry_bytes(P)), u8(L.gsb_image_bytes(self.W, self.H))
        self.binning = u8(L.gsb_binning_bytes(self.max_rendered))
        a = RasterArgs()
        a.P, a.D, a.M, a.width, a.height = P, 0, 0, self.W, self.H
        a.background, a.means3D, a.colors_precomp = self.bg.data_ptr(), self.means_cam.data_ptr(), self.params.ptr("rgb")
        a.opacities, a.scales, a.scale_modifier, a.rotations = self.opac.data_ptr(), self.scales.data_ptr(), 1.0, self.rot.data_ptr()
        a.viewmatrix, a.projmatrix, a.cam_pos = self.view.data_ptr(), self.proj.data_ptr(), self.campos.data_ptr()
        a.tan_fovx, a.tan_fovy = self.tanfovx, self.tanfovy
        self.args = a
        sp = self.g_side.data_ptr()
        self.gout = GradOutputs(dL_dmean2D=sp, dL_dconic=sp + 12 * P, dL_dcov3D=sp + 28 * P, dL_dopacity=self.g_opac.data_ptr(),
                                dL_dcolor=self.grads.ptr("rgb"), dL_dmean3D=self.g_means_cam.data_ptr(), dL_dsh=None,
                                dL_dscale=self.g_scales.data_ptr(), dL_drot=self.g_rot.data_ptr())

    def save_ply(self, path: str) -> None:
        """Write the map as the reference's GaussianModel.ply (src/Utils.cc:182-280; read by scripts/replay.py)."""
        from .ply import save_gaussian_model
        c = lambda name: self.params[name].detach().cpu().numpy()
        save_gaussian_model(path, c("means"), c("rgb"), c("opacity"), c("scales"), c("quats"))

    @classmethod
    def from_ply(cls, path: str, **kwargs) -> "MapOptimizer":
        """Resume from a GaussianModel.ply written by the reference or by ``save_ply``."""
        from .ply import load_gaussian_model
        m = load_gaussian_model(path)
        return cls(m["means"], m["rgb"], m["logit_opacities"], m["log_scales"], m["unnorm_quats"], **kwargs)

    def _s(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def render(self, Tcw: torch.Tensor):
        """Forward only: prologue + rasterizer.  Returns (color [3,H,W], depth [1,H,W], radii [P])."""
        L, p, s = self.L, self.params, self._s()
        self._Tcw = Tcw.to(self.dev, torch.float32).contiguous()
        with torch.cuda.device(self.dev):
            _lib.check(L.gsb_prologue(self.P, self._Tcw.data_ptr(), p.ptr("means"), p.ptr("opacity"), p.ptr("quats"), p.ptr("scales"),
                                      self.means_cam.data_ptr(), self.opac.data_ptr(), self.rot.data_ptr(), self.scales.data_ptr(), s))
            _lib.check(L.gsb_forward_ws(C.byref(self.args), self.geom.data_ptr(), self.geom.numel(), self.binning.data_ptr(),
                                        self.binning.numel(), self.max_rendered, self.img.data_ptr(), self.img.numel(),
                                        self.color.data_ptr(), self.depth.data_ptr(), self.radii.data_ptr(), s))
        return self.color, self.depth, self.radii

    def backward(self, dL_dpix: torch.Tensor):
        """Rasterizer backward + prologue backward into the packed gradient block (and dL/dTcw)."""
        L, p, g, s = self.L, self.params, self.grads, self._s()
        dL = dL_dpix.to(self.dev, torch.float32).contiguous()
        with torch.cuda.device(self.dev):
            _lib.check(L.gsb_backward(C.byref(self.args), -1, self.radii.data_ptr(), self.geom.data_ptr(), self.binning.data_ptr(),
                                      self.img.data_ptr(), dL.data_ptr(), C.byref(self.gout), s))
            _lib.check(L.gsb_prologue_backward(self.P, self._Tcw.data_ptr(), p.ptr("means"), p.ptr("opacity"), p.ptr("quats"),
                                               p.ptr("scales"), self.g_means_cam.data_ptr(), self.g_opac.data_ptr(),
                                               self.g_rot.data_ptr(), self.g_scales.data_ptr(), g.ptr("means"), g.ptr("opacity"),
                                               g.ptr("quats"), g.ptr("scales"), self.dTcw.data_ptr(), s))
        return g

    def render_fused(self, Tcw: torch.Tensor):
        """Prologue + ONE five-channel rasterization: (color [3,H,W], depth_sil [2,H,W], median_depth [1,H,W], radii)
        -- what Render::RenderForFrame gets from its depth pass and its RGB pass (src/Render.cc:445-448)."""
        L, p, s = self.L, self.params, self._s()
        self._Tcw = Tcw.to(self.dev, torch.float32).contiguous()
        if not hasattr(self, "depth_sil"):
            self.depth_sil = torch.empty((2, self.H, self.W), dtype=torch.float32, device=self.dev)
            self.g_z = torch.empty(self.P, dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(L.gsb_prologue(self.P, self._Tcw.data_ptr(), p.ptr("means"), p.ptr("opacity"), p.ptr("quats"), p.ptr("scales"),
                                      self.means_cam.data_ptr(), self.opac.data_ptr(), self.rot.data_ptr(), self.scales.data_ptr(), s))
            _lib.check(L.gsb_forward_fused_ws(C.byref(self.args), self.geom.data_ptr(), self.geom.numel(), self.binning.data_ptr(),
                                              self.binning.numel(), self.max_rendered, self.img.data_ptr(), self.img.numel(),
                                              self.color.data_ptr(), self.depth_sil.data_ptr(), self.depth.data_ptr(),
                                              self.radii.data_ptr(), s))
        return self.color, self.depth_sil, self.depth, self.radii

    def backward_fused(self, dL_dcolor: torch.Tensor, dL_ddepth_sil: torch.Tensor, z_attached: bool = True):
        """Backward of ``render_fused`` into the packed gradient block.  ``z_attached``: the depth pass' z_cam colour is a
        function of the means (mapping mode, src/Render.cc:973-976); False = tracking mode (detached, :957)."""
        L, p, g, s = self.L, self.params, self.grads, self._s()
        dC = dL_dcolor.to(self.dev, torch.float32).contiguous()
        dD = dL_ddepth_sil.to(self.dev, torch.float32).contiguous()
        with torch.cuda.device(self.dev):
            _lib.check(L.gsb_backward_fused(C.byref(self.args), self.radii.data_ptr(), self.geom.data_ptr(), self.binning.data_ptr(),
                                            self.img.data_ptr(), dC.data_ptr(), dD.data_ptr(), C.byref(self.gout), self.g_z.data_ptr(),
                                            1 if z_attached else 0, s))
            _lib.check(L.gsb_prologue_backward(self.P, self._Tcw.data_ptr(), p.ptr("means"), p.ptr("opacity"), p.ptr("quats"),
                                               p.ptr("scales"), self.g_means_cam.data_ptr(), self.g_opac.data_ptr(),
                                               self.g_rot.data_ptr(), self.g_scales.data_ptr(), g.ptr("means"), g.ptr("opacity"),
                                               g.ptr("quats"), g.ptr("scales"), self.dTcw.data_ptr(), s))
        return g

    def adam(self):
        """torch::optim::Adam step on every group, reading the (all-reduced) gradient block in place: ONE launch over the
        packed [14, P] block with a learning rate per group (gsb_adam_step_groups)."""
        self.t += 1
        L, s = self.L, self._s()
        if not hasattr(self, "_adam_sizes"):
            self._adam_sizes = (C.c_longlong * len(GROUPS))(*[w * self.P for _, w in GROUPS])
        lrs = (C.c_float * len(GROUPS))(*[float(self.lr[name]) for name, _ in GROUPS])
        with torch.cuda.device(self.dev):
            _lib.check(L.gsb_adam_step_groups(len(GROUPS), self._adam_sizes, lrs, self.params.flat.data_ptr(), self.grads.flat.data_ptr(),
                                              self.exp_avg.flat.data_ptr(), self.exp_avg_sq.flat.data_ptr(), float(self.betas[0]),
                                              float(self.betas[1]), self.eps, self.t, s))

    def _exchange_and_adam(self, average: bool):
        if self.exchange is not None:
            self.exchange.allreduce(self.grads.flat, use_multicast=world()[1] >= 4)
            if average:
                self.grads.flat.mul_(1.0 / world()[1])
        else:
            allreduce_gradients(self.grads, average=average)
        self.adam()

    def step_fused(self, Tcw: torch.Tensor, loss_grad: Callable, average: bool = False, z_attached: bool = True):
        """One optimisation step with ONE rasterization.  ``loss_grad(color, depth_sil, median_depth) ->
        (dL/dcolor [3,H,W], dL/ddepth_sil [2,H,W])``."""
        color, depth_sil, median, _ = self.render_fused(Tcw)
        dC, dD = loss_grad(color, depth_sil, median)
        self.backward_fused(dC, dD, z_attached=z_attached)
        self._exchange_and_adam(average)
        return color, depth_sil

    def step_slam(self, Tcw: torch.Tensor, gt_color: torch.Tensor, gt_depth: torch.Tensor, lambda_: float = 0.8,
                  w_image: float = 1.0, w_depth: float = 0.7, w_surdepth: float = 0.35, average: bool = False):
        """One complete mapping iteration of Render::RenderForFrame (src/Render.cc:420-476) without a single torch op on the
        hot path: prologue -> ONE five-channel rasterization -> fused L1 + SSIM + depth loss and its gradient
        (gsb_mapping_loss) -> summed backward -> prologue backward -> exchange -> Adam.  Weights default to
        Examples/RGB-D/replica.yaml:89-94.  Returns the 8 loss terms (device tensor: l1, ssim, depth_l1, surdepth_l1, total,
        n_valid, n_valid_sur, 0); the scale regularisers (Render.cc:462-467) are the caller's."""
        L, s = self.L, self._s()
        color, depth_sil, median, _ = self.render_fused(Tcw)
        if not hasattr(self, "_loss_scratch"):
            nb = int(L.gsb_loss_scratch_bytes(self.W, self.H))
            self._loss_scratch = torch.empty(nb, dtype=torch.uint8, device=self.dev)
            self._gC = torch.empty((3, self.H, self.W), dtype=torch.float32, device=self.dev)
            self._gD = torch.empty((2, self.H, self.W), dtype=torch.float32, device=self.dev)
            self.loss_terms = torch.empty(8, dtype=torch.float32, device=self.dev)
        gtc = gt_color.to(self.dev, torch.float32).contiguous()
        gtd = gt_depth.to(self.dev, torch.float32).contiguous()
        with torch.cuda.device(self.dev):
            _lib.check(L.gsb_mapping_loss(self.W, self.H, color.data_ptr(), depth_sil.data_ptr(), median.data_ptr(), gtc.data_ptr(),
                                          gtd.data_ptr(), float(lambda_), float(w_image), float(w_depth), float(w_surdepth),
                                          self._gC.data_ptr(), self._gD.data_ptr(), self.loss_terms.data_ptr(),
                                          self._loss_scratch.data_ptr(), self._loss_scratch.numel(), s))
        self.backward_fused(self._gC, self._gD, z_attached=True)
        self._exchange_and_adam(average)
        return self.loss_terms

    def step(self, Tcw: torch.Tensor, loss_grad: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], average: bool = False):
        """One optimisation step on this rank's keyframe.  ``loss_grad(color, depth) -> dL/dcolor``."""
        color, depth, _ = self.render(Tcw)
        self.backward(loss_grad(color, depth))
        self._exchange_and_adam(average)
        return color
