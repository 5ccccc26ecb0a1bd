"""Camera-pose refinement through the rasterizer -- the loop of ``Render::RenderStartTraking`` (src/Render.cc:985-1141) over
the C ABI: the Gaussians are fixed, the 7 pose parameters (unnormalised quaternion + translation, ``Gaussian::InitCameraPose``,
src/Gaussian.cc:98-128) follow the gradient of a masked L1 image / depth loss.

Per iteration: ``Tcw = Rt2T(q, t)`` (src/Utils.cc:170-179, ``ToRotation`` include/Utils.h:56-77) -> ONE five-channel
rasterization with the depth colours detached (tracking mode, src/Render.cc:957) -> masked L1 sums over the pixels whose
silhouette exceeds 0.99 (:1075-1092) -> rasterizer backward -> ``gsb_prologue_backward`` reduces ``dL/dTcw = sum_i g_i [p_i;1]^T``
on the device (the reference materialises an N x 4 x 4 repeat + bmm for it) -> the 12 numbers are chained to (q, t) in closed form
on the host -> Adam with the reference's learning rates (both groups use the quaternion rate, src/Gaussian.cc:149-150).
The ORB reprojection term (:1012-1065): matched map points X_w, undistorted keypoints and their inverse level variances come
from the ORB front end (host data, ``set_features``); the term sum_inliers e^T diag(1/sigma^2) e, e = K X_c / X_c.z - obs, is a
few hundred points and is differentiated in closed form on the host; the chi-square gate (5.991, two degrees of
freedom) is applied once, at half the iteration budget (:1081-1085); ``run`` stops early when the loss moves by less than 1e-3
(:1109-1110) and returns the best pose seen (:1101-1108).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

from .mapping import MapOptimizer


def rt2T(q: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """``Rt2T`` / ``ToRotation``: [R(q / |q|) | t; 0 0 0 1] for q = (w, x, y, z) of shape [4], t of shape [3]."""
    qn = q / torch.sqrt((q * q).sum())
    r0, x, y, z = qn[0], qn[1], qn[2], qn[3]
    R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r0 * z), 2 * (x * z + r0 * y)]),
                     torch.stack([2 * (x * y + r0 * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r0 * x)]),
                     torch.stack([2 * (x * z - r0 * y), 2 * (y * z + r0 * x), 1 - 2 * (x * x + y * y)])])
    top = torch.cat([R, t.reshape(3, 1)], 1)
    return torch.cat([top, torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=q.device, dtype=q.dtype)], 0)


def rt2T_np(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """``rt2T`` on the host: [4,4] float32.  Scalar double arithmetic rounded once (the tracking loop runs this, the chain rule below
    and Adam between two GPU iterations with the device idle: plain Python floats cost about 10 us, a numpy expression of the same 23 + 37 + 25 us)."""
    w, x, y, z = q.tolist()
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    t0, t1, t2 = t.tolist()
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), t0],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x), t1],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y), t2],
                     [0.0, 0.0, 0.0, 1.0]], dtype=np.float32)


def rt2T_backward_np(q: np.ndarray, G: np.ndarray):
    """dL/dq [4], dL/dt [3] from G = dL/dTcw (rows 0..2 of the 4x4 matrix, [3,4]): the autograd of ``rt2T`` in closed form
    (dR/dw, dR/dx, dR/dy, dR/dz contracted with G[:, :3], then through q / |q|)."""
    w, x, y, z = q.tolist()
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    (g00, g01, g02, g03), (g10, g11, g12, g13), (g20, g21, g22, g23) = G[:3].tolist()
    gw = 2.0 * (-z * g01 + y * g02 + z * g10 - x * g12 - y * g20 + x * g21)
    gx = 2.0 * (y * g01 + z * g02 + y * g10 - 2 * x * g11 - w * g12 + z * g20 + w * g21 - 2 * x * g22)
    gy = 2.0 * (-2 * y * g00 + x * g01 + w * g02 + x * g10 + z * g12 - w * g20 + z * g21 - 2 * y * g22)
    gz = 2.0 * (-2 * z * g00 - w * g01 + x * g02 + w * g10 - 2 * z * g11 + y * g12 + x * g20 + y * g21)
    dot = w * gw + x * gx + y * gy + z * gz
    gq = np.array([(gw - w * dot) / n, (gx - x * dot) / n, (gy - y * dot) / n, (gz - z * dot) / n], dtype=np.float32)
    return gq, np.array([g03, g13, g23], dtype=np.float32)


class PoseOptimizer:
    """State on the HOST: the seven pose parameters, their Adam moments and the ORB matches are a few hundred floats, and the
    loop reads the loss back every iteration anyway (the reference's ``loss.item()``), so the chain rule through ``Rt2T``, the
    reprojection term and the Adam update are closed-form numpy (about 0.1 ms) instead of ~100 small torch kernels per iteration
    (1.4 ms of launch overhead); the rasterizer, the masked L1 loss and dL/dTcw stay on the device."""

    def __init__(self, gaussians: MapOptimizer, quat, trans, lr_quat: float = 2e-3, betas=(0.9, 0.999), eps: float = 1e-15,
                 use_graph: bool = True):
        self.g = gaussians
        self.use_graph, self._graph, self._graph_key = bool(use_graph), None, None
        as_np = lambda a: (a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)).astype(np.float32).copy()
        self._q, self._t = as_np(quat).reshape(4), as_np(trans).reshape(3)
        # CreateOptimizerForPose: the translation group is created with the QUATERNION learning rate (src/Gaussian.cc:150)
        self.lr, self.betas, self.eps = float(lr_quat), (float(betas[0]), float(betas[1])), float(eps)
        self._adam = dict(step=0, m=np.zeros(7, np.float32), v=np.zeros(7, np.float32))
        self.best = (self._q.copy(), self._t.copy(), float("inf"))
        self._rendered_once = False   # the first step runs the whole prologue (the map may have changed since the last frame)
        self.features = None
        self.last_terms = {}

    # torch views of the state (tests, callers that hand the pose on)
    @property
    def q(self) -> torch.Tensor:
        return torch.from_numpy(self._q.copy()).to(self.g.dev)

    @property
    def t(self) -> torch.Tensor:
        return torch.from_numpy(self._t.copy()).to(self.g.dev)

    def pose(self) -> torch.Tensor:
        return torch.from_numpy(rt2T_np(self._q, self._t)).to(self.g.dev)

    def set_features(self, K, Xw, obs, inv_sigma2) -> None:
        """ORB matches of the frame (src/Render.cc:1012-1047): ``K`` [3,3] intrinsics, ``Xw`` [M,3] matched map points,
        ``obs`` [M,2] undistorted keypoints, ``inv_sigma2`` [M] inverse variance of each keypoint's pyramid level."""
        f32 = lambda a: (a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)).astype(np.float32)
        self.features = dict(K=f32(K), Xw=f32(Xw), obs=f32(obs), w=f32(inv_sigma2).reshape(-1),
                             inlier=np.ones(f32(Xw).shape[0], dtype=bool))

    def _reprojection_np(self, T: np.ndarray, with_grad: bool = False):
        f = self.features
        Xc = f["Xw"] @ T[:3, :3].T + T[:3, 3]
        iz = 1.0 / Xc[:, 2:3]
        u = Xc * iz
        e = (u @ f["K"].T)[:, :2] - f["obs"]
        err = (e * e).sum(1) * f["w"]
        if not with_grad:
            return err
        ge = 2.0 * f["w"][:, None] * e * f["inlier"][:, None]                     # d(sum of inlier errors) / d e
        gu = ge @ f["K"][:2, :]                                                   # [M,3]; u_z is constant
        gXc = np.stack([gu[:, 0] * iz[:, 0], gu[:, 1] * iz[:, 0], -(gu[:, 0] * u[:, 0] + gu[:, 1] * u[:, 1]) * iz[:, 0]], 1)
        G = np.zeros((3, 4), np.float32)
        G[:, :3] = gXc.T @ f["Xw"]
        G[:, 3] = gXc.sum(0)
        return err, G

    def reprojection(self, Tcw: torch.Tensor):
        """Weighted squared reprojection error of every match, [M] (src/Render.cc:1058-1065)."""
        return torch.from_numpy(self._reprojection_np(Tcw.detach().cpu().numpy().astype(np.float32))).to(self.g.dev)

    def _adam_step(self, grad: np.ndarray) -> None:
        """torch::optim::Adam on the 7 parameters (bias-corrected, eps outside the square root; src/Gaussian.cc:145-175)."""
        a = self._adam
        a["step"] += 1
        b1, b2 = self.betas
        c1, c2 = 1.0 - b1 ** a["step"], 1.0 - b2 ** a["step"]
        step_size, rc2 = self.lr / c1, 1.0 / math.sqrt(c2)
        m, v, p = a["m"].tolist(), a["v"].tolist(), self._q.tolist() + self._t.tolist()
        for k, g in enumerate(grad.tolist()):
            m[k] = b1 * m[k] + (1 - b1) * g
            v[k] = b2 * v[k] + (1 - b2) * g * g
            p[k] -= step_size * m[k] / (math.sqrt(v[k]) * rc2 + self.eps)
        a["m"], a["v"] = np.array(m, dtype=np.float32), np.array(v, dtype=np.float32)
        self._q, self._t = np.array(p[:4], dtype=np.float32), np.array(p[4:], dtype=np.float32)

    def step(self, gt_color: torch.Tensor, gt_depth: torch.Tensor, w_image: float = 1.0, w_depth: float = 1.0,
             use_surdepth: bool = True, w_feature: float = 0.0, gate_features: bool = False) -> float:
        """One iteration; returns the loss.  ``use_surdepth``: the depth term reads the median depth, which carries no
        gradient (include/Rasterizer.cuh:210), exactly as with ``Tracking.useSurDepth: true`` in the shipped YAMLs.
        ``w_feature`` > 0 (and ``set_features`` called) adds the ORB reprojection term; ``gate_features`` re-selects the
        inliers (chi-square 5.991) before it is summed -- the reference does that once, at half its iteration budget."""
        g = self.g
        T = rt2T_np(self._q, self._t)
        if not hasattr(self, "_dC"):
            self._dC = torch.empty((3, g.H, g.W), dtype=torch.float32, device=g.dev)
            self._dD = torch.empty((2, g.H, g.W), dtype=torch.float32, device=g.dev)
            self._out = torch.zeros(20, dtype=torch.float32, device=g.dev)        # loss terms [8] | dL/dTcw [12]
            self._host = torch.empty(20, dtype=torch.float32).pin_memory()
            self._T_host = torch.empty((4, 4), dtype=torch.float32).pin_memory()
            self._T_dev = torch.empty((4, 4), dtype=torch.float32, device=g.dev)
        dC, dD, terms = self._dC, self._dD, self._out[:8]
        gtc, gtd = gt_color.contiguous(), gt_depth.contiguous()
        # The whole iteration is queued before the host waits for anything (one synchronisation per iteration, the reference's
        # loss.item()): pose upload from pinned memory -> prologue (after this optimizer's first render only the camera-frame means:
        # the map is frozen while the pose is tracked) -> five-channel pass -> masked L1 terms and their gradients in ONE kernel
        # (gsb_tracking_loss: mask = "uncertainDepth", src/Render.cc:1075; sums over the mask, L1LossForTracking, src/Utils.cc:45-52)
        # -> per-pixel backward + pose-only per-Gaussian kernel (dL/dTcw straight into the read-back block) -> D2H.  The forward's
        # overflow latch is read after that; an overflowing frame is redone with a larger binning blob.
        self._T_host.copy_(torch.from_numpy(T))
        stream = torch.cuda.current_stream(g.dev)

        def enqueue(in_graph: bool):
            self._T_dev.copy_(self._T_host, non_blocking=True)
            g._Tcw = self._T_dev
            g._forward_fused(means_only=self._rendered_once, record_event=not in_graph)
            with torch.cuda.device(g.dev):
                _lib.check(g.L.gsb_tracking_loss(g.W, g.H, g.color.data_ptr(), g.depth_sil.data_ptr(), g.depth.data_ptr(), gtc.data_ptr(),
                                                 gtd.data_ptr(), float(w_image), float(w_depth), 1 if use_surdepth else 0,
                                                 dC.data_ptr(), dD.data_ptr(), terms.data_ptr(), torch.cuda.current_stream(g.dev).cuda_stream))
            g.backward_pose(dC, dD, z_attached=False, out=self._out[8:])
            self._host.copy_(self._out, non_blocking=True)

        while True:
            # From the second iteration on the device side of the iteration is a fixed sequence over fixed buffers (only the
            # CONTENT of the pinned pose buffer changes): it is captured once as a CUDA graph and replayed -- one launch per
            # iteration instead of about twenty, and no gaps between the kernels.  The key covers everything that is baked in.
            key = (gtc.data_ptr(), gtd.data_ptr(), float(w_image), float(w_depth), bool(use_surdepth), g.P, g.max_rendered,
                   g.binning.data_ptr(), g.params.flat.data_ptr())
            if self.use_graph and self._rendered_once:
                if self._graph_key != key:
                    stream.synchronize()
                    self._graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self._graph):
                        enqueue(True)
                    self._graph_key = key
                self._graph.replay()
                stream.synchronize()
                overflowed = g._overflowed(wait=False)
            else:
                enqueue(False)
                stream.synchronize()
                overflowed = g._overflowed()
            self._rendered_once = True
            if not overflowed:
                break
        h = self._host.numpy()
        image_term, depth_term, loss = float(h[0]), float(h[1]), float(h[2])
        G = h[8:].reshape(3, 4).astype(np.float32).copy()
        feat = 0.0
        if w_feature > 0.0 and self.features is not None and self.features["Xw"].shape[0] > 0:
            if gate_features:
                self.features["inlier"] = self._reprojection_np(T) < 5.991        # src/Render.cc:1081-1084
            err, Gf = self._reprojection_np(T, with_grad=True)
            feat = float(err[self.features["inlier"]].sum())
            loss += w_feature * feat
            G += np.float32(w_feature) * Gf
        self.last_terms = dict(image=image_term, depth=depth_term, feature=feat)
        if loss == loss and loss < self.best[2]:                                   # best-so-far bookkeeping, :1101-1108
            self.best = (self._q.copy(), self._t.copy(), loss)
        gq, gt = rt2T_backward_np(self._q, G)
        self._adam_step(np.concatenate([gq, gt]))
        return loss

    def run(self, gt_color: torch.Tensor, gt_depth: torch.Tensor, iters: int = 200, w_image: float = 0.7, w_depth: float = 1.0,
            w_feature: float = 0.1, use_surdepth: bool = True, tol: float = 1e-3):
        """``Render::RenderStartTraking``'s loop (src/Render.cc:1052-1127) with the weights of Examples/RGB-D/tum/TUM1.yaml:
        up to ``iters`` iterations, chi-square gate of the ORB matches at iteration iters / 2, early exit when the loss moves by
        less than ``tol`` (the update of that iteration is NOT applied, as in the reference), best pose kept.  Returns
        (best Tcw [4,4], best loss, iterations run)."""
        gate_at = int(iters / 2.0)
        last = 0.0
        n = 0
        for it in range(iters):
            q0, t0 = self._q.copy(), self._t.copy()
            state = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self._adam.items()}
            loss = self.step(gt_color, gt_depth, w_image, w_depth, use_surdepth, w_feature, gate_features=(it == gate_at))
            n = it + 1
            if abs(last - loss) < tol:
                # the reference breaks BEFORE StepUpdataForPose (:1109-1110): undo this iteration's Adam step
                self._q, self._t, self._adam = q0, t0, state
                break
            last = loss
        bq, bt, bl = self.best
        return torch.from_numpy(rt2T_np(bq, bt)).to(self.g.dev), bl, n
