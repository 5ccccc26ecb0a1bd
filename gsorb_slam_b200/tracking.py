"""Camera-pose refinement through the rasterizer -- the loop of ``Render::RenderStartTraking`` (src/Render.cc:985-1141) over
the C ABI: the Gaussians are fixed, the 7 pose parameters (unnormalised quaternion + translation, ``Gaussian::InitCameraPose``,
src/Gaussian.cc:98-128) follow the gradient of a masked L1 image / depth loss.

Per iteration: ``Tcw = Rt2T(q, t)`` (src/Utils.cc:170-179, ``ToRotation`` include/Utils.h:56-77) -> ONE five-channel
rasterization with the depth colours detached (tracking mode, src/Render.cc:957) -> masked L1 sums over the pixels whose
silhouette exceeds 0.99 (:1075-1092) -> rasterizer backward -> ``gsb_prologue_backward`` reduces ``dL/dTcw = sum_i g_i [p_i;1]^T``
on the device (the reference materialises an N x 4 x 4 repeat + bmm for it) -> the 12 numbers are chained to (q, t) by autograd
on a 4x4 matrix -> Adam with the reference's learning rates (both groups use the quaternion rate, src/Gaussian.cc:149-150).
The ORB reprojection term (:1012-1065): matched map points X_w, undistorted keypoints and their inverse level variances come
from the ORB front end (host data, ``set_features``); the term sum_inliers e^T diag(1/sigma^2) e, e = K X_c / X_c.z - obs, is a
few hundred points and is differentiated by torch autograd on the same 4x4 matrix; the chi-square gate (5.991, two degrees of
freedom) is applied once, at half the iteration budget (:1081-1085); ``run`` stops early when the loss moves by less than 1e-3
(:1109-1110) and returns the best pose seen (:1101-1108).
"""
from __future__ import annotations

import torch

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


class PoseOptimizer:
    def __init__(self, gaussians: MapOptimizer, quat, trans, lr_quat: float = 2e-3, betas=(0.9, 0.999), eps: float = 1e-15):
        self.g = gaussians
        dev = gaussians.dev
        self.q = torch.as_tensor(quat, dtype=torch.float32, device=dev).clone().requires_grad_(True)
        self.t = torch.as_tensor(trans, dtype=torch.float32, device=dev).clone().requires_grad_(True)
        # CreateOptimizerForPose: the translation group is created with the QUATERNION learning rate (src/Gaussian.cc:150)
        self.adam = torch.optim.Adam([{"params": [self.q], "lr": lr_quat}, {"params": [self.t], "lr": lr_quat}], betas=betas, eps=eps)
        self.best = (self.q.detach().clone(), self.t.detach().clone(), float("inf"))
        self.features = None
        self.last_terms = {}

    def pose(self) -> torch.Tensor:
        return rt2T(self.q, self.t)

    def set_features(self, K, Xw, obs, inv_sigma2) -> None:
        """ORB matches of the frame (src/Render.cc:1012-1047): ``K`` [3,3] intrinsics, ``Xw`` [M,3] matched map points,
        ``obs`` [M,2] undistorted keypoints, ``inv_sigma2`` [M] inverse variance of each keypoint's pyramid level."""
        dev = self.g.dev
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)
        self.features = dict(K=f32(K), Xw=f32(Xw), obs=f32(obs), w=f32(inv_sigma2).reshape(-1),
                             inlier=torch.ones(f32(Xw).shape[0], dtype=torch.bool, device=dev))

    def reprojection(self, Tcw: torch.Tensor):
        """Weighted squared reprojection error of every match, [M] (src/Render.cc:1058-1065)."""
        f = self.features
        Xc = f["Xw"] @ Tcw[:3, :3].T + Tcw[:3, 3]
        uv = (Xc / Xc[:, 2:3]) @ f["K"].T
        e = uv[:, :2] - f["obs"]
        return (e * e).sum(1) * f["w"]

    def step(self, gt_color: torch.Tensor, gt_depth: torch.Tensor, w_image: float = 1.0, w_depth: float = 1.0,
             use_surdepth: bool = True, w_feature: float = 0.0, gate_features: bool = False) -> float:
        """One iteration; returns the loss.  ``use_surdepth``: the depth term reads the median depth, which carries no
        gradient (include/Rasterizer.cuh:210), exactly as with ``Tracking.useSurDepth: true`` in the shipped YAMLs.
        ``w_feature`` > 0 (and ``set_features`` called) adds the ORB reprojection term; ``gate_features`` re-selects the
        inliers (chi-square 5.991) before it is summed -- the reference does that once, at half its iteration budget."""
        g = self.g
        Tcw = self.pose()
        color, depth_sil, median, _ = g.render_fused(Tcw.detach())
        mask = (depth_sil[1] > 0.99) & ~torch.isnan(gt_depth)                      # "uncertainDepth", src/Render.cc:1075
        dI = color - gt_color
        dC = (w_image * torch.sign(dI) * mask).contiguous()
        dD = torch.zeros_like(depth_sil)
        if use_surdepth:
            depth_term = (median[0] - gt_depth).abs()[mask].sum()
        else:
            dd = depth_sil[0] - gt_depth
            dD[0] = w_depth * torch.sign(dd) * mask
            depth_term = dd.abs()[mask].sum()
        image_term = dI.abs()[mask.expand_as(dI)].sum()
        feat = None
        if w_feature > 0.0 and self.features is not None and self.features["Xw"].shape[0] > 0:
            err = self.reprojection(Tcw)
            if gate_features:
                self.features["inlier"] = (err < 5.991).detach()                   # src/Render.cc:1081-1084
            feat = err[self.features["inlier"]].sum()
        loss = float(w_image * image_term + w_depth * depth_term + (w_feature * feat.detach() if feat is not None else 0.0))
        self.last_terms = dict(image=float(image_term), depth=float(depth_term), feature=float(feat.detach()) if feat is not None else 0.0)
        g.backward_fused(dC, dD, z_attached=False)                                 # -> g.dTcw [3,4] on the device
        self.adam.zero_grad()
        grad = torch.zeros(4, 4, device=g.dev)
        grad[:3] = g.dTcw
        Tcw.backward(gradient=grad, retain_graph=feat is not None)
        if feat is not None:
            (w_feature * feat).backward()
        if loss == loss and loss < self.best[2]:                                   # best-so-far bookkeeping, :1101-1108
            self.best = (self.q.detach().clone(), self.t.detach().clone(), loss)
        self.adam.step()
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
            q0, t0 = self.q.detach().clone(), self.t.detach().clone()
            state = {k: {kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} for k, v in self.adam.state.items()}
            loss = self.step(gt_color, gt_depth, w_image, w_depth, use_surdepth, w_feature, gate_features=(it == gate_at))
            n = it + 1
            if abs(last - loss) < tol:
                # the reference breaks BEFORE StepUpdataForPose (:1109-1110): undo this iteration's Adam step
                with torch.no_grad():
                    self.q.copy_(q0); self.t.copy_(t0)
                for k, v in state.items():
                    self.adam.state[k] = v
                break
            last = loss
        bq, bt, bl = self.best
        with torch.no_grad():
            T = rt2T(bq, bt)
        return T, bl, n
