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
from ._lib import GradOutputs, MapUpdate, RasterArgs
from .distributed import BLOCK_ROWS, GROUPS, GradBlock, SymmetricExchange, allreduce_gradients, world

# Examples/RGB-D/replica.yaml:95-99 (Mapping.lrs*)
DEFAULT_LR = {"means": 1e-4, "rgb": 2.5e-3, "quats": 1e-3, "opacity": 0.05, "scales": 1e-3}


def keyframe_batch_hyperparameters(G: int, lr: Optional[Dict[str, float]] = None, betas=(0.9, 0.999)):
    """Optimiser settings for the G-rank keyframe-batch shard (one G-view minibatch Adam step instead of G single-view steps,
    SURVEY.md 8e): gradients AVERAGED over the ranks, learning rates x G / 2, betas -> betas ** G (the moment averages then
    forget over the same number of FRAMES).  Measured by tools/minibatch_parity.py (profiles/r02_minibatch_parity.json, 200 k
    Gaussians, 16 training / 4 held-out views, 480 frames): held-out PSNR +0.014 dB against the reference's sequential
    schedule at G = 8 (x G instead of x G / 2: +0.19 dB); unscaled settings lose 4.1 dB."""
    lr = dict(DEFAULT_LR if lr is None else lr)
    return {k: v * max(1.0, G / 2.0) for k, v in lr.items()}, (betas[0] ** G, betas[1] ** G)


def backproject_pixels(L, dev, W: int, H: int, mask, gt_depth, gt_color, Tcw, fx: float, fy: float, cx: float, cy: float,
                       max_z: Optional[torch.Tensor] = None):
    """``Render::ProjectPixel`` / ``Render::InitGaussianPoint`` + the SinglePixel initialisation of
    ``Gaussian::AddGaussianPoints`` (src/Render.cc:617-655, 666-707; src/Gaussian.cc:50-74) on the device
    (``gsb_backproject``): one Gaussian per pixel with ``mask >= 250`` (``mask`` None: every pixel) and depth > 0, in
    row-major pixel order.  Returns (means [K,3], rgb [K,3], logit_opacities [K], log_scales [K,3], unnorm_quats [K,4]);
    ``max_z`` (1 device float, the running ``Render::mMaxZ``) is raised to the largest selected depth."""
    d = torch.device(dev)
    gtc = gt_color.to(d, torch.float32).contiguous()
    gtd = gt_depth.to(d, torch.float32).contiguous()
    Twc = torch.linalg.inv(Tcw.detach().to("cpu", torch.float64)).to(torch.float32).contiguous()
    cap = int(W) * int(H)
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=d)
    means, rgb, ls, quat, op = e(cap, 3), e(cap, 3), e(cap, 3), e(cap, 4), e(cap)
    count = torch.zeros(1, dtype=torch.int32, device=d)
    nb = int(L.gsb_backproject_scratch_bytes(int(W), int(H)))
    scratch = torch.empty(nb, dtype=torch.uint8, device=d)
    Th = (C.c_float * 16)(*Twc.reshape(-1).tolist())
    with torch.cuda.device(d):
        _lib.check(L.gsb_backproject(int(W), int(H), mask.data_ptr() if mask is not None else None, gtd.data_ptr(), gtc.data_ptr(),
                                     float(fx), float(fy), float(cx), float(cy), Th, cap, means.data_ptr(), rgb.data_ptr(),
                                     ls.data_ptr(), quat.data_ptr(), op.data_ptr(), count.data_ptr(),
                                     max_z.data_ptr() if max_z is not None else None, scratch.data_ptr(), nb,
                                     torch.cuda.current_stream(d).cuda_stream))
    K = int(count.item())
    return means[:K], rgb[:K], op[:K], ls[:K], quat[:K]


class MapOptimizer:
    """The Gaussian map of one rank and its optimiser state, resident on the GPU as ARENAS (``GradBlock`` with a capacity):
    parameters, gradients and both Adam moments can grow in place (``add_gaussians``, ``densify``) and shrink
    (``prune_low_opacity``) without a ``torch::cat`` / ``index_select`` per tensor (src/Gaussian.cc:209-258)."""

    def __init__(self, means, rgb, logit_opacities, log_scales, unnorm_quats, *, width, height, tanfovx, tanfovy,
                 projmatrix, background=None, lr: Optional[Dict[str, float]] = None, betas=(0.9, 0.999), eps=1e-15,
                 device="cuda:0", max_rendered: Optional[int] = None, capacity: Optional[int] = None, scene_radius: float = 0.0,
                 w_reg_scalar: float = 10.0, w_reg_long: float = 5.0):
        self.L = _lib.lib()
        self.dev = torch.device(device)
        d = self.dev
        t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(d, torch.float32)
        self.P = int(t(means).shape[0])
        P = self.P
        # an automatic capacity is a multiple of 4 rows: every group of the arenas then starts 16-byte aligned and the fused
        # update (gsb_backward_fused_update) stages its rows with bulk copies; an explicit capacity is taken as given
        self.capacity = max(int(capacity), P) if capacity else (P + 3) // 4 * 4
        cap = self.capacity
        self.params = GradBlock(P, d, capacity=cap)   # same layout as the gradient block
        self.params["means"].copy_(t(means)); self.params["rgb"].copy_(t(rgb))
        self.params["opacity"].copy_(t(logit_opacities).reshape(P, 1)); self.params["scales"].copy_(t(log_scales))
        self.params["quats"].copy_(t(unnorm_quats))
        # world size > 1: the gradient block lives in a symmetric allocation and is all-reduced by ONE libgsb kernel
        # over NVLink (distributed.SymmetricExchange); NCCL only if peer mapping is unavailable
        self.exchange = None
        if world()[1] > 1 and self.dev.type == "cuda":
            try:
                self.exchange = SymmetricExchange(BLOCK_ROWS * cap, d)
            except Exception:
                self.exchange = None
        self.grads = GradBlock(P, d, storage=self.exchange.alloc(BLOCK_ROWS * cap) if self.exchange else None, capacity=cap)
        self.exp_avg, self.exp_avg_sq = GradBlock(P, d, capacity=cap), GradBlock(P, d, capacity=cap)
        self.lr = dict(DEFAULT_LR if lr is None else lr)
        self.betas, self.eps, self.t = betas, float(eps), 0
        self.W, self.H = int(width), int(height)
        self.tanfovx, self.tanfovy = float(tanfovx), float(tanfovy)
        self.proj = t(projmatrix).reshape(16).contiguous()
        # default mode of Render::StartSplatting: identity view, means pre-transformed (src/Render.cc:748-754)
        self.view = torch.eye(4, device=d).reshape(16).contiguous()
        self.campos = torch.zeros(3, device=d)
        self.bg = t(background if background is not None else np.zeros(3, np.float32))
        # scale regularisers of the mapping loss (src/Render.cc:418, :462-469): maxScalar = 0.1 * scene radius
        # (Render::mMaxZ / sceneRaduisDepthRatio, :661); weights Examples/RGB-D/replica.yaml:93-94.  0 = regularisers off
        self.scene_radius, self.w_reg_scalar, self.w_reg_long = float(scene_radius), float(w_reg_scalar), float(w_reg_long)
        self.dTcw = torch.empty((3, 4), dtype=torch.float32, device=d)
        self.color = torch.empty((3, self.H, self.W), dtype=torch.float32, device=d)
        self.depth = torch.empty((1, self.H, self.W), dtype=torch.float32, device=d)
        self.reg_terms = torch.zeros(8, dtype=torch.float32, device=d)
        self.max_rendered = int(max_rendered) if max_rendered else 4 * cap + 4096
        self.img = torch.empty(int(self.L.gsb_image_bytes(self.W, self.H)), dtype=torch.uint8, device=d)
        self.binning = torch.empty(int(self.L.gsb_binning_bytes(self.max_rendered)), dtype=torch.uint8, device=d)
        # forward status (GeomHeader words: magic, P, num_rendered, binned, overflow latch, ...) polled once per step
        self._status = torch.zeros(8, dtype=torch.int32).pin_memory() if d.type == "cuda" else torch.zeros(8, dtype=torch.int32)
        self._status_ev = torch.cuda.Event() if d.type == "cuda" else None
        self.overflow_retries = 0
        self._alloc_rows()

    # ---- per-row temporaries and the C-ABI argument blocks: rebuilt when the arena grows, re-pointed when P changes ----
    def _alloc_rows(self):
        d, cap = self.dev, self.capacity
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=d)
        # activated temporaries + their gradients
        self.means_cam, self.opac, self.rot, self.scales = e(cap, 3), e(cap), e(cap, 4), e(cap, 3)
        self.g_means_cam, self.g_opac, self.g_rot, self.g_scales = e(cap, 3), e(cap), e(cap, 4), e(cap, 3)
        self.g_side = e(cap, 13)   # mean2D 3 | conic 4 | cov3D 6: reference outputs nobody consumes here
        self.radii = torch.empty(cap, dtype=torch.int32, device=d)
        self.geom = torch.empty(int(self.L.gsb_geometry_bytes(cap)), dtype=torch.uint8, device=d)
        if hasattr(self, "g_z"):
            self.g_z = e(cap)
        self._point_args()

    def _point_args(self):
        P, cap = self.P, self.capacity
        a = RasterArgs()
        a.P, a.D, a.M, a.width, a.height = P, 0, 0, self.W, self.H
        a.background, a.means3D, a.colors_precomp = self.bg.data_ptr(), self.means_cam.data_ptr(), self.params.ptr("rgb")
        a.opacities, a.scales, a.scale_modifier, a.rotations = self.opac.data_ptr(), self.scales.data_ptr(), 1.0, self.rot.data_ptr()
        a.viewmatrix, a.projmatrix, a.cam_pos = self.view.data_ptr(), self.proj.data_ptr(), self.campos.data_ptr()
        a.tan_fovx, a.tan_fovy = self.tanfovx, self.tanfovy
        self.args = a
        sp = self.g_side.data_ptr()
        self.gout = GradOutputs(dL_dmean2D=sp, dL_dconic=sp + 12 * cap, dL_dcov3D=sp + 28 * cap, dL_dopacity=self.g_opac.data_ptr(),
                                dL_dcolor=self.grads.ptr("rgb"), dL_dmean3D=self.g_means_cam.data_ptr(), dL_dsh=None,
                                dL_dscale=self.g_scales.data_ptr(), dL_drot=self.g_rot.data_ptr())
        self._adam_sizes = (C.c_longlong * len(GROUPS))(*self.params.group_sizes())

    # ---- growing and shrinking the map --------------------------------------------------------------------------------
    def reserve(self, capacity: int) -> None:
        """Grow the arenas to ``capacity`` rows (parameters, gradients, Adam moments, per-row temporaries)."""
        if capacity <= self.capacity:
            return
        if self.exchange is not None:
            raise RuntimeError("the gradient block of a multi-GPU map lives in a symmetric allocation: create the MapOptimizer "
                               "with capacity=... large enough for the session")
        self.params, self.exp_avg, self.exp_avg_sq = self.params.grown(capacity), self.exp_avg.grown(capacity), self.exp_avg_sq.grown(capacity)
        self.grads = GradBlock(self.P, self.dev, capacity=capacity)
        self.capacity = int(capacity)
        self._alloc_rows()

    def add_gaussians(self, means, rgb, logit_opacities, log_scales, unnorm_quats) -> int:
        """``Gaussian::AddGaussianPoints`` -> ``UpdateOptimizerParams`` -> ``CatTensorToOptimizer`` (src/Gaussian.cc:50-95,
        241-258): K new rows behind the existing ones, zero Adam moments, the step counter is shared.  In place while the
        arena has room; otherwise the arena grows by half (at least to fit)."""
        K = int(means.shape[0])
        if K == 0:
            return 0
        P = self.P
        if P + K > self.capacity:
            self.reserve((max(P + K, self.capacity + self.capacity // 2) + 3) // 4 * 4)   # groups stay 16-byte aligned
        for blk in (self.params, self.grads, self.exp_avg, self.exp_avg_sq):
            blk.resize(P + K)
        new = dict(means=means, rgb=rgb, opacity=logit_opacities.reshape(K, 1), scales=log_scales, quats=unnorm_quats)
        for name, _ in GROUPS:
            self.params[name][P:].copy_(new[name].to(self.dev, torch.float32).reshape(self.params[name][P:].shape))
            self.exp_avg[name][P:].zero_()
            self.exp_avg_sq[name][P:].zero_()
            self.grads[name][P:].zero_()
        self.P = P + K
        if 4 * self.P + 4096 > self.max_rendered:
            self._grow_binning(4 * self.capacity + 4096)
        self._point_args()
        return K

    def add_mask(self, color, depth_sil, gt_depth, median_mul: float = 0.5):
        """The densification mask of ``Render::AddGaussian`` (src/Render.cc:557-583) on the device: pixels that are dark, not
        yet opaque and off in depth by more than an adaptive threshold (mean + median_mul * median of the small depth errors,
        at least 1 cm), or whose silhouette is below 0.8.  Returns a uint8 [H, W] mask (255 = add), as ``ImshowDepth`` hands
        it to ``ProjectPixel`` (which keeps values >= 250)."""
        gray = (color[0] * 299 + color[1] * 587 + color[2] * 114) / 1000
        black = gray < 50 / 255.0
        diff = (gt_depth - depth_sil[0]).abs()
        small = (diff < 0.05) & (gt_depth > 0) & (depth_sil[0] > 0)
        sel = diff[small]
        th = float(sel.sum() / small.sum() + median_mul * sel.median()) if sel.numel() else float("nan")
        if not th >= 0.01:   # also NaN (empty selection)
            th = 0.01
        c1 = ~(depth_sil[1] > 0.99) & black & (diff > th)
        c2 = depth_sil[1] < 0.8
        return ((c1 | c2).to(torch.uint8) * 255).contiguous()

    def densify(self, Tcw, gt_color, gt_depth, fx: float, fy: float, cx: float, cy: float, color=None, depth_sil=None,
                median_mul: float = 0.5, mask=None) -> int:
        """``Render::AddGaussian`` for one keyframe: densification mask (from the last ``render_fused`` outputs unless given)
        -> GPU back-projection of the selected pixels (``gsb_backproject`` = ProjectPixel + the SinglePixel initialisation of
        AddGaussianPoints) -> append.  Returns the number of Gaussians added; updates ``scene_radius`` bookkeeping input
        ``max_z``."""
        d = self.dev
        gtd = gt_depth.to(d, torch.float32).contiguous()
        if mask is None:
            color = self.color if color is None else color
            depth_sil = self.depth_sil if depth_sil is None else depth_sil
            mask = self.add_mask(color, depth_sil, gtd, median_mul)
        if not hasattr(self, "max_z"):
            self.max_z = torch.zeros(1, dtype=torch.float32, device=d)
        rows = backproject_pixels(self.L, d, self.W, self.H, mask, gtd, gt_color, Tcw, fx, fy, cx, cy, self.max_z)
        return self.add_gaussians(rows[0], rows[1], rows[2], rows[3], rows[4])

    def prune_low_opacity(self, threshold: float = 0.005) -> int:
        """``Render::RemoveGaussian`` (src/Render.cc:598-616; Gaussian.cc:193-239, pruneOpcities 0.005): rows whose
        sigmoid(logit opacity) is below the threshold leave the parameters and both Adam moments, order preserved --
        one flag kernel and ONE compaction pass over the 15 tensors (gsb_low_opacity_keep + gsb_prune_rows)."""
        L, d, P = self.L, self.dev, self.P
        keep = torch.empty(P, dtype=torch.uint8, device=d)
        src_blocks = (self.params, self.exp_avg, self.exp_avg_sq)
        dst_blocks = tuple(GradBlock(P, d, capacity=self.capacity) for _ in src_blocks)
        n = 3 * len(GROUPS)
        src = (C.c_void_p * n)(*[blk.ptr(name) for blk in src_blocks for name, _ in GROUPS])
        dst = (C.c_void_p * n)(*[blk.ptr(name) for blk in dst_blocks for name, _ in GROUPS])
        widths = (C.c_int * n)(*[w for _ in src_blocks for _, w in GROUPS])
        count = torch.zeros(1, dtype=torch.int32, device=d)
        nb = int(L.gsb_prune_scratch_bytes(P))
        scratch = torch.empty(nb, dtype=torch.uint8, device=d)
        with torch.cuda.device(d):
            _lib.check(L.gsb_low_opacity_keep(P, self.params.ptr("opacity"), float(threshold), keep.data_ptr(), self._s()))
            _lib.check(L.gsb_prune_rows(P, keep.data_ptr(), n, src, dst, widths, count.data_ptr(), scratch.data_ptr(), nb, self._s()))
        K = int(count.item())
        if K == P:
            return 0
        self.params, self.exp_avg, self.exp_avg_sq = dst_blocks
        for blk in (self.params, self.exp_avg, self.exp_avg_sq, self.grads):
            blk.resize(K)
        if self.exchange is None:
            self.grads.flat.zero_()
        self.P = K
        self._point_args()
        return P - K

    # ---- the two loops around the iteration: Render::InitWorld and Render::RenderForFrame ------------------------------------
    @classmethod
    def init_world(cls, Tcw, gt_color, gt_depth, fx: float, fy: float, cx: float, cy: float, *, iters: int = 200,
                   radius_depth_ratio: float = 3.0, lambda_: float = 0.8, w_image: float = 1.0, w_depth: float = 0.7,
                   w_surdepth: float = 0.1, device="cuda:0", **kwargs) -> "MapOptimizer":
        """``Render::InitWorld`` (src/Render.cc:496-553): the first RGB-D frame becomes the map -- one Gaussian per pixel with a
        valid depth, back-projected at the frame's pose in raster order (``InitGaussianPoint``, :666-707) -- and is fitted to
        that frame for ``iters`` iterations (200 in the reference) with the image + depth loss of :523-531: no scale
        regulariser, the median-depth term weighted 0.1 (it carries no gradient, include/Rasterizer.cuh:210, so only the reported
        loss depends on it).  ``scene_radius`` = max depth / ``radius_depth_ratio`` (:705, ``Mapping.raduisDepthRatio``) is set
        on the returned optimizer for the mapping iterations that follow.  ``kwargs``: width, height, tanfovx, tanfovy,
        projmatrix, lr ... of the constructor."""
        d = torch.device(device)
        L = _lib.lib()
        W, H = int(kwargs["width"]), int(kwargs["height"])
        max_z = torch.zeros(1, dtype=torch.float32, device=d)
        Tcw = Tcw.to(d, torch.float32)
        rows = backproject_pixels(L, d, W, H, None, gt_depth, gt_color, Tcw, fx, fy, cx, cy, max_z)
        if rows[0].shape[0] == 0:
            raise ValueError("InitWorld: the frame has no pixel with a valid depth")
        kwargs.pop("scene_radius", None)
        mo = cls(rows[0], rows[1], rows[2], rows[3], rows[4], device=d, scene_radius=0.0, **kwargs)
        mo.max_z = max_z
        gtc, gtd = gt_color.to(d, torch.float32).contiguous(), gt_depth.to(d, torch.float32).contiguous()
        for _ in range(int(iters)):
            mo.step_slam(Tcw, gtc, gtd, lambda_, w_image, w_depth, w_surdepth, average=world()[1] > 1)   # every rank holds the same frame
        mo.scene_radius = float(max_z.item()) / float(radius_depth_ratio)
        return mo

    def update_scene_radius(self, radius_depth_ratio: float = 3.0) -> float:
        """``mSceneRadius = mMaxZ / raduisDepthRatio`` (src/Render.cc:661) after a densification raised ``max_z``."""
        if hasattr(self, "max_z"):
            self.scene_radius = float(self.max_z.item()) / float(radius_depth_ratio)
        return self.scene_radius

    def map_keyframes(self, keyframes, iters: int = 60, rng=None, **loss_weights):
        """The loop of ``Render::RenderForFrame`` (src/Render.cc:402-493): ``iters`` mapping iterations (``Mapping.numIters``, 60
        in replica.yaml), each on ONE keyframe drawn uniformly from the candidate window (:423) -- ``keyframes`` is that window,
        a sequence of (Tcw [4,4], gt_color [3,H,W], gt_depth [H,W]) which the caller selected (covisibility lives in the ORB
        front end).  With G ranks every iteration draws G keyframes from the same generator and rank r takes the r-th: the
        keyframe-batch shard (gradients averaged over the ranks; see ``keyframe_batch_hyperparameters``).  ``rng``: a
        ``random.Random`` (seeded alike on every rank).  Returns the loss terms of the last iteration (device tensor)."""
        import random
        rng = rng if rng is not None else random.Random(0)
        rank, G = world()
        kf = [(T.to(self.dev, torch.float32).contiguous(), c.to(self.dev, torch.float32).contiguous(),
               z.to(self.dev, torch.float32).contiguous()) for T, c, z in keyframes]
        if not kf:
            raise ValueError("map_keyframes: empty keyframe window")
        terms = None
        for _ in range(int(iters)):
            draws = [rng.randrange(len(kf)) for _ in range(G)]
            T, c, z = kf[draws[rank]]
            terms = self.step_slam(T, c, z, average=G > 1, **loss_weights)
        return terms

    def _grow_binning(self, max_rendered: int) -> None:
        self.max_rendered = int(max_rendered)
        self.binning = torch.empty(int(self.L.gsb_binning_bytes(self.max_rendered)), dtype=torch.uint8, device=self.dev)

    # ---- the forward's overflow latch: polled once per step, never ignored ---------------------------------------------
    def _poll_forward(self, record_event: bool = True):
        """Stream-ordered copy of the forward's status words to pinned host memory (no synchronisation here).  Inside a CUDA
        graph capture no event is recorded (an event recorded during capture cannot be waited for from the host)."""
        self._status.copy_(self.geom[:32].view(torch.int32), non_blocking=True)
        if record_event:
            self._status_ev.record(torch.cuda.current_stream(self.dev))

    def _overflowed(self, wait: bool = True) -> bool:
        """True if the last forward found more tile instances than ``max_rendered``: the binning blob is grown and the caller
        renders again (what adapter/Rasterizer.cc does for the drop-in path).  Waits for the FORWARD only (``wait=False``: the
        caller has synchronised the stream already)."""
        if wait:
            self._status_ev.synchronize()
        if int(self._status[4]) == 0:
            return False
        need = int(self._status[2]) & 0xffffffff
        self._grow_binning(max(need + need // 4 + 4096, 2 * self.max_rendered))
        self.overflow_retries += 1
        return True

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
        """Forward only: prologue + rasterizer.  Returns (color [3,H,W], depth [1,H,W], radii [P]).  A frame with more tile
        instances than the binning blob holds is rendered again with a larger blob (never silently truncated)."""
        L, p = self.L, self.params
        self._Tcw = Tcw.to(self.dev, torch.float32).contiguous()
        while True:
            s = self._s()
            with torch.cuda.device(self.dev):
                _lib.check(L.gsb_prologue(self.P, self._Tcw.data_ptr(), p.ptr("means"), p.ptr("opacity"), p.ptr("quats"), p.ptr("scales"),
                                          self.means_cam.data_ptr(), self.opac.data_ptr(), self.rot.data_ptr(), self.scales.data_ptr(), s))
                _lib.check(L.gsb_forward_ws(C.byref(self.args), self.geom.data_ptr(), self.geom.numel(), self.binning.data_ptr(),
                                            self.binning.numel(), self.max_rendered, self.img.data_ptr(), self.img.numel(),
                                            self.color.data_ptr(), self.depth.data_ptr(), self.radii.data_ptr(), s))
                self._poll_forward()
            if not self._overflowed():
                break
        return self.color, self.depth, self.radii[:self.P]

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

    def _forward_fused(self, means_only: bool = False, record_event: bool = True):
        """``means_only``: the map has not changed since the last prologue (tracking): only the camera-frame means are redone."""
        L, p, s = self.L, self.params, self._s()
        if not hasattr(self, "depth_sil"):
            self.depth_sil = torch.empty((2, self.H, self.W), dtype=torch.float32, device=self.dev)
            self.g_z = torch.empty(self.capacity, dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            act = (None, None, None) if means_only else (self.opac.data_ptr(), self.rot.data_ptr(), self.scales.data_ptr())
            _lib.check(L.gsb_prologue(self.P, self._Tcw.data_ptr(), p.ptr("means"), p.ptr("opacity"), p.ptr("quats"), p.ptr("scales"),
                                      self.means_cam.data_ptr(), *act, s))
            _lib.check(L.gsb_forward_fused_ws(C.byref(self.args), self.geom.data_ptr(), self.geom.numel(), self.binning.data_ptr(),
                                              self.binning.numel(), self.max_rendered, self.img.data_ptr(), self.img.numel(),
                                              self.color.data_ptr(), self.depth_sil.data_ptr(), self.depth.data_ptr(),
                                              self.radii.data_ptr(), s))
            self._poll_forward(record_event)

    def render_fused(self, Tcw: torch.Tensor, frozen_map: bool = False):
        """Prologue + ONE five-channel rasterization: (color [3,H,W], depth_sil [2,H,W], median_depth [1,H,W], radii)
        -- what Render::RenderForFrame gets from its depth pass and its RGB pass (src/Render.cc:445-448).  Stand-alone use:
        an overflowing frame is rendered again with a larger binning blob before this returns.  ``frozen_map``: the caller
        guarantees that no parameter changed since the previous render of this optimizer (the tracking loop): the activations
        of opacity / rotation / scale are reused and only the camera-frame means are recomputed."""
        self._Tcw = Tcw.to(self.dev, torch.float32).contiguous()
        while True:
            self._forward_fused(means_only=frozen_map)
            if not self._overflowed():
                break
        return self.color, self.depth_sil, self.depth, self.radii[:self.P]

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

    def backward_pose(self, dL_dcolor: torch.Tensor, dL_ddepth_sil: torch.Tensor, z_attached: bool = False,
                      out: Optional[torch.Tensor] = None):
        """Backward of ``render_fused`` as far as the camera pose: ``dTcw`` [3,4] and nothing else (gsb_backward_fused_pose) --
        the tracking loop's backward (src/Render.cc:1052-1127 optimises the pose only).  ``out``: 12 device floats that receive
        the gradient instead of ``dTcw``."""
        L, s = self.L, self._s()
        dC = dL_dcolor.to(self.dev, torch.float32).contiguous()
        dD = dL_ddepth_sil.to(self.dev, torch.float32).contiguous()
        with torch.cuda.device(self.dev):
            _lib.check(L.gsb_backward_fused_pose(C.byref(self.args), self.radii.data_ptr(), self.geom.data_ptr(), self.binning.data_ptr(),
                                                 self.img.data_ptr(), dC.data_ptr(), dD.data_ptr(), 1 if z_attached else 0,
                                                 self.params.ptr("means"), (self.dTcw if out is None else out).data_ptr(), s))
        return self.dTcw if out is None else out

    def add_scale_regularisers(self):
        """reg_scalar / reg_long of the mapping loss (src/Render.cc:462-469) added to the log-scale gradients; no-op while
        ``scene_radius`` is 0.  ``reg_terms`` (device) = {reg_scalar, reg_long, selected (row, axis) pairs, 0}."""
        if not self.scene_radius > 0.0:
            return
        with torch.cuda.device(self.dev):
            _lib.check(self.L.gsb_scale_regulariser(self.P, self.params.ptr("scales"), 0.1 * self.scene_radius, self.w_reg_scalar,
                                                    self.w_reg_long, self.grads.ptr("scales"), self.reg_terms.data_ptr(), self._s()))

    def adam(self):
        """torch::optim::Adam step on every group, reading the (all-reduced) gradient block in place: ONE launch over the
        whole arena with a learning rate per group (gsb_adam_step_groups; padding rows are zero and stay zero)."""
        self.t += 1
        L, s = self.L, self._s()
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
        self.add_scale_regularisers()
        self._exchange_and_adam(average)
        return color, depth_sil

    def slam_gradients(self, Tcw: torch.Tensor, gt_color: torch.Tensor, gt_depth: torch.Tensor, lambda_: float = 0.8,
                       w_image: float = 1.0, w_depth: float = 0.7, w_surdepth: float = 0.35, _update: bool = False,
                       _write_grads: bool = False):
        """The gradient half of a mapping iteration (src/Render.cc:420-470) without a single torch op on the hot path:
        prologue -> ONE five-channel rasterization -> fused L1 + SSIM + depth loss and its gradient (gsb_mapping_loss) ->
        summed backward -> prologue backward, into ``grads``.  The forward's overflow latch is read once (the host waits for
        the FORWARD only, with the loss and the backward already queued behind it): an overflowing frame is redone with a
        larger binning blob.  Returns the 8 loss terms (device tensor: l1, ssim, depth_l1, surdepth_l1, total of the pixel
        terms, n_valid, n_valid_sur, 0)."""
        L = self.L
        self._Tcw = Tcw.to(self.dev, torch.float32).contiguous()
        if not hasattr(self, "_loss_scratch"):
            nb = int(L.gsb_loss_scratch_bytes(self.W, self.H))
            self._loss_scratch = torch.empty(nb, dtype=torch.uint8, device=self.dev)
            self._gC = torch.empty((3, self.H, self.W), dtype=torch.float32, device=self.dev)
            self._gD = torch.empty((2, self.H, self.W), dtype=torch.float32, device=self.dev)
            self.loss_terms = torch.empty(8, dtype=torch.float32, device=self.dev)
        gtc = gt_color.to(self.dev, torch.float32).contiguous()
        gtd = gt_depth.to(self.dev, torch.float32).contiguous()
        while True:
            self._forward_fused()
            with torch.cuda.device(self.dev):
                _lib.check(L.gsb_mapping_loss(self.W, self.H, self.color.data_ptr(), self.depth_sil.data_ptr(), self.depth.data_ptr(),
                                              gtc.data_ptr(), gtd.data_ptr(), float(lambda_), float(w_image), float(w_depth),
                                              float(w_surdepth), self._gC.data_ptr(), self._gD.data_ptr(), self.loss_terms.data_ptr(),
                                              self._loss_scratch.data_ptr(), self._loss_scratch.numel(), self._s()))
                if _update:   # an overflowed forward leaves the map untouched (the kernel reads the latch): safe to queue
                    _lib.check(L.gsb_backward_fused_update(C.byref(self.args), self.radii.data_ptr(), self.geom.data_ptr(),
                                                           self.binning.data_ptr(), self.img.data_ptr(), self._gC.data_ptr(),
                                                           self._gD.data_ptr(), 1, C.byref(self._map_update(_write_grads)), self._s()))
            if not _update:
                self.backward_fused(self._gC, self._gD, z_attached=True)
            if not self._overflowed():
                break
        return self.loss_terms

    def _map_update(self, write_grads: bool) -> MapUpdate:
        """gsb_map_update for Adam step ``t + 1`` over the five arenas."""
        u = MapUpdate()
        u.Tcw = self._Tcw.data_ptr()
        for g, (name, _) in enumerate(GROUPS):
            u.params[g], u.exp_avg[g], u.exp_avg_sq[g] = self.params.ptr(name), self.exp_avg.ptr(name), self.exp_avg_sq.ptr(name)
            u.grads[g] = self.grads.ptr(name) if write_grads else None
            u.lr[g] = float(self.lr[name])
        u.beta1, u.beta2, u.eps, u.step = float(self.betas[0]), float(self.betas[1]), self.eps, self.t + 1
        u.dL_dTcw = self.dTcw.data_ptr()
        u.max_scalar = 0.1 * self.scene_radius if self.scene_radius > 0.0 else 0.0
        u.w_scalar, u.w_long, u.reg_terms = self.w_reg_scalar, self.w_reg_long, self.reg_terms.data_ptr()
        return u

    def step_slam(self, Tcw: torch.Tensor, gt_color: torch.Tensor, gt_depth: torch.Tensor, lambda_: float = 0.8,
                  w_image: float = 1.0, w_depth: float = 0.7, w_surdepth: float = 0.35, average: bool = False,
                  fused_update: Optional[bool] = None, write_grads: bool = False):
        """One complete mapping iteration of Render::RenderForFrame (src/Render.cc:420-476).  Weights default to
        Examples/RGB-D/replica.yaml:89-94.
        One rank (``fused_update`` None or True): everything behind the per-pixel backward -- per-Gaussian backward, chain
        rule of the prologue, scale regularisers (when ``scene_radius`` is set; ``reg_terms``), Adam -- is ONE launch
        (gsb_backward_fused_update) and no gradient array is written unless ``write_grads``.
        Several ranks, or ``fused_update=False``: ``slam_gradients`` -> regularisers -> exchange -> Adam as separate passes
        (the gradient block has to exist between them)."""
        single = world()[1] == 1
        if fused_update is None:
            fused_update = single
        if fused_update and not single:
            raise ValueError("the fused update skips the gradient exchange: single rank only")
        if fused_update:
            terms = self.slam_gradients(Tcw, gt_color, gt_depth, lambda_, w_image, w_depth, w_surdepth, _update=True,
                                        _write_grads=write_grads)
            self.t += 1
            return terms
        terms = self.slam_gradients(Tcw, gt_color, gt_depth, lambda_, w_image, w_depth, w_surdepth)
        self.add_scale_regularisers()
        self._exchange_and_adam(average)
        return terms

    def step(self, Tcw: torch.Tensor, loss_grad: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], average: bool = False):
        """One optimisation step on this rank's keyframe.  ``loss_grad(color, depth) -> dL/dcolor``."""
        color, depth, _ = self.render(Tcw)
        self.backward(loss_grad(color, depth))
        self.add_scale_regularisers()
        self._exchange_and_adam(average)
        return color
