"""Camera-pose refinement through the rasterizer -- the loop of ``Render::RenderStartTraking`` (src/Render.cc:985-1141) over
the C ABI: the Gaussians are fixed, the 7 pose parameters (unnormalised quaternion + translation, ``Gaussian::InitCameraPose``,
src/Gaussian.cc:98-128) follow the gradient of a masked L1 image / depth loss.

Per iteration: ``Tcw = Rt2T(q, t)`` (src/Utils.cc:170-179, ``ToRotation`` include/Utils.h:56-77) -> ONE five-channel
rasterization with the depth colours detached (tracking mode, src/Render.cc:957) -> masked L1 sums over the pixels whose
silhouette exceeds 0.99 (:1075-1092) -> rasterizer backward -> ``gsb_prologue_backward`` reduces ``dL/dTcw = sum_i g_i [p_i;1]^T``
on the device (the reference materialises an N x 4 x 4 repeat + bmm for it) -> the 12 numbers are chained to (q, t) by autograd
on a 4x4 matrix -> Adam with the reference's learning rates (both groups use the quaternion rate, src/Gaussian.cc:149-150).
The ORB reprojection term of the reference (:1058-1065, :1081-1085) is a host-side input there and is not part of this helper.
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

    def pose(self) -> torch.Tensor:
        return rt2T(self.q, self.t)

    def step(self, gt_color: torch.Tensor, gt_depth: torch.Tensor, w_image: float = 1.0, w_depth: float = 1.0,
             use_surdepth: bool = True) -> float:
        """One iteration; returns the loss.  ``use_surdepth``: the depth term reads the median depth, which carries no
        gradient (include/Rasterizer.cuh:210), exactly as with ``Tracking.useSurDepth: true`` in the shipped YAMLs."""
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
        loss = float(w_image * dI.abs()[mask.expand_as(dI)].sum() + w_depth * depth_term)
        g.backward_fused(dC, dD, z_attached=False)                                 # -> g.dTcw [3,4] on the device
        self.adam.zero_grad()
        grad = torch.zeros(4, 4, device=g.dev)
        grad[:3] = g.dTcw
        Tcw.backward(gradient=grad)
        if loss == loss and loss < self.best[2]:                                   # best-so-far bookkeeping, :1101-1108
            self.best = (self.q.detach().clone(), self.t.detach().clone(), loss)
        self.adam.step()
        return loss
