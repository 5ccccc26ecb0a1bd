"""A complete RGB-D session on the library's host-side mirror of ``Render`` -- the loops of the reference around the rasterizer:

    frame 0            Render::InitWorld            (src/Render.cc:496-553)   MapOptimizer.init_world
    every later frame  Render::RenderStartTraking   (src/Render.cc:985-1141)  PoseOptimizer.run            (pose of the new frame)
                       Render::AddGaussian          (src/Render.cc:557-594)   MapOptimizer.densify         (grow the map)
                       Render::RenderForFrame       (src/Render.cc:402-493)   MapOptimizer.map_keyframes   (60 iterations on the window)
                       Render::RemoveGaussian       (src/Render.cc:598-616)   MapOptimizer.prune_low_opacity
    at the end         SaveGaussianModel            (src/Utils.cc:182-280)    MapOptimizer.save_ply / from_ply

There is no dataset on the GPU box, so the "sensor" is a hidden ground-truth map (a textured, gently curved wall) rendered by the
library itself from a short camera trajectory; the SLAM side never sees it, only its RGB-D frames.  The ORB front end (key points,
covisibility) is outside this library: the tracking loop runs without the reprojection term and the keyframe window is simply the
last few frames.

    python examples/slam_session.py [--frames 5] [--width 160 --height 120] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gsorb_slam_b200.mapping import MapOptimizer   # noqa: E402
from gsorb_slam_b200.scene import Camera           # noqa: E402
from gsorb_slam_b200.tracking import PoseOptimizer, rt2T_np   # noqa: E402


def psnr_metric(img: torch.Tensor, gt: torch.Tensor) -> float:
    """``PSNRMetric`` (src/Utils.cc:33-37): mean over the channels of 20 log10(1 / rmse)."""
    mse = ((img - gt) ** 2).reshape(img.shape[0], -1).mean(1)
    return float((20.0 * torch.log10(1.0 / mse.clamp_min(1e-20).sqrt())).mean())


def wall_world(cam: Camera, margin: float = 0.25, seed: int = 0):
    """The hidden scene: one Gaussian per pixel of a frame ``margin`` wider than the camera's on every side, on the surface
    z(u, v) = 3 + 0.4 sin(u / 25) + 0.3 cos(v / 18), smooth colours with a little noise, nearly opaque, sigma = 1.2 pixels."""
    W, H, fx, fy = cam.width, cam.height, cam.fx, cam.fy
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(-margin * W, (1 + margin) * W, 1.0), np.arange(-margin * H, (1 + margin) * H, 1.0))
    u, v = u.reshape(-1).astype(np.float32), v.reshape(-1).astype(np.float32)
    z = (3.0 + 0.4 * np.sin(u / 25.0) + 0.3 * np.cos(v / 18.0)).astype(np.float32)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    means = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1).astype(np.float32)
    rgb = np.stack([0.5 + 0.4 * np.sin(u / 9.0 + ph) * np.cos(v / 7.0 - ph) for ph in (0.0, 1.3, 2.9)], 1)
    rgb = np.clip(rgb + rng.normal(0, 0.03, rgb.shape), 0.05, 0.95).astype(np.float32)
    n = means.shape[0]
    log_scales = np.repeat(np.log(1.2 * z / ((fx + fy) / 2.0))[:, None], 3, 1).astype(np.float32)
    quats = np.zeros((n, 4), np.float32)
    quats[:, 0] = 1.0
    return means, rgb, np.full(n, 4.0, np.float32), log_scales, quats


def trajectory(n: int):
    """Camera poses Tcw of the frames: a slow pan (1 cm .. 3 cm and 0.6 degrees per frame)."""
    poses = []
    for k in range(n):
        a = 0.01 * k
        q = np.array([np.cos(a / 2), 0.0, np.sin(a / 2), 0.0], np.float32)   # rotation about y
        t = np.array([0.03 * k, -0.01 * k, 0.02 * k], np.float32)
        poses.append((q, t))
    return poses


def run(width: int = 160, height: int = 120, frames: int = 5, init_iters: int = 60, track_iters: int = 60, map_iters: int = 30,
        window: int = 3, device: str = "cuda:0", seed: int = 0, ply_path: str | None = None) -> dict:
    dev = torch.device(device)
    cam = Camera(width, height, 0.8 * width, 0.8 * width)
    fx, fy, cx, cy = cam.fx, cam.fy, (width - 1) / 2.0, (height - 1) / 2.0
    kw = dict(width=width, height=height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, projmatrix=cam.projmatrix)
    # ---- the sensor ----
    world = MapOptimizer(*wall_world(cam, seed=seed), device=dev, **kw)
    poses = trajectory(frames)
    sensor = []
    for q, t in poses:
        T = torch.from_numpy(rt2T_np(q, t)).to(dev)
        color, depth_sil, median, _ = world.render_fused(T)
        depth = torch.where(depth_sil[1] > 0.5, median[0], torch.zeros_like(median[0]))   # 0 = no measurement
        sensor.append((T, color.clamp(0, 1).clone(), depth.clone()))
    del world
    out = {"image": f"{width}x{height}", "frames": frames, "per_frame": []}
    t0 = time.time()
    # ---- frame 0: InitWorld ----
    T0, c0, d0 = sensor[0]
    slam = MapOptimizer.init_world(T0, c0, d0, fx, fy, cx, cy, iters=0, device=dev, **kw)
    out["init_gaussians"] = slam.P
    out["init_valid_pixels"] = int((d0 > 0).sum())
    out["psnr_frame0_before_init_db"] = psnr_metric(slam.render_fused(T0)[0].clamp(0, 1), c0)
    first = last = None
    for it in range(init_iters):   # InitWorld's loop (the classmethod runs it when iters > 0; spelled out here to log the loss)
        terms = slam.step_slam(T0, c0, d0, w_surdepth=0.1)
        if it == 0:
            first = float(terms[4])
        last = float(terms[4])
    slam.update_scene_radius()
    out["init_loss_first_last"] = [first, last]
    out["psnr_frame0_after_init_db"] = psnr_metric(slam.render_fused(T0)[0].clamp(0, 1), c0)
    out["scene_radius"] = slam.scene_radius
    # ---- the session ----
    rng = random.Random(seed)
    q_est, t_est = poses[0]
    keyframes = [sensor[0]]
    for k in range(1, frames):
        T_true, c, d = sensor[k]
        # tracking: the pose of the previous frame is the starting point (Gaussian::InitCameraPose, src/Gaussian.cc:98-128)
        po = PoseOptimizer(slam, q_est, t_est)
        err0 = float((po.pose() - T_true).abs().max())
        T_est, best_loss, n_it = po.run(c, d, iters=track_iters, w_image=0.7, w_depth=1.0, w_feature=0.0)
        bq, bt, _ = po.best
        q_est, t_est = bq.copy(), bt.copy()
        err1 = float((T_est - T_true).abs().max())
        # mapping: densify where this frame is not explained yet, then optimise over the window, then prune
        slam.render_fused(T_est)
        added = slam.densify(T_est, c, d, fx, fy, cx, cy)
        slam.update_scene_radius()
        keyframes = (keyframes + [(T_est, c, d)])[-window:]
        terms = slam.map_keyframes(keyframes, iters=map_iters, rng=rng)
        removed = slam.prune_low_opacity(0.005)
        out["per_frame"].append({"frame": k, "pose_err_before": err0, "pose_err_after": err1, "tracking_iters": n_it,
                                 "tracking_loss": best_loss, "added": added, "removed": removed, "gaussians": slam.P,
                                 "mapping_loss": float(terms[4]),
                                 "psnr_db": psnr_metric(slam.render_fused(T_est)[0].clamp(0, 1), c)})
    torch.cuda.synchronize(dev)
    out["seconds"] = time.time() - t0
    out["overflow_retries"] = slam.overflow_retries
    # ---- GaussianModel.ply round trip ----
    path = ply_path or os.path.join(tempfile.mkdtemp(), "GaussianModel.ply")
    slam.save_ply(path)
    again = MapOptimizer.from_ply(path, device=dev, **kw)
    a, b = slam.render_fused(sensor[-1][0])[0].clone(), again.render_fused(sensor[-1][0])[0]
    out["ply_round_trip_bit_identical"] = bool(torch.equal(a, b))
    out["ply_round_trip_max_abs_diff"] = float((a - b).abs().max())
    out["final_gaussians"] = slam.P
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--height", type=int, default=120)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    res = run(a.width, a.height, a.frames)
    line = json.dumps(res)
    print(line)
    if a.json:
        with open(a.json, "w") as f:
            f.write(line + "\n")
