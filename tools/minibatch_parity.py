#!/usr/bin/env python
"""Statistical parity of the keyframe-batch shard.

The reference's mapping loop draws ONE keyframe per Adam step (src/Render.cc:420-476).  The G-GPU keyframe-batch shard turns
that into G-view minibatch steps: rank r renders keyframe r, the per-Gaussian gradients are summed by the exchange step, every
rank takes the same Adam step.  That is a different optimiser trajectory, so parity is statistical: PSNR (src/Utils.cc:33-37:
10 log10(1 / mse)) on held-out views after the SAME frames have been consumed.

Experiment (synthetic, one GPU; a sum of G per-frame gradients is what gsb_exchange_allreduce delivers, bit for bit, so the
G-rank job is emulated by accumulating G backward passes before one Adam step):
  * ground truth: a seeded SLAM-like map, rendered from 16 training poses and 4 held-out poses by the library itself;
  * start: the ground-truth map with its parameters perturbed;
  * schedule A (the reference): 60 G sequential single-view Adam steps over a fixed random frame sequence;
  * schedule B (the shard): 60 Adam steps, step i uses frames [G i, G i + G) of the SAME sequence, gradients summed or averaged,
    learning rates scaled by a rule.
Prints one JSON object; `python tools/minibatch_parity.py --P 200000 --G 8`.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsorb_slam_b200.mapping import DEFAULT_LR, MapOptimizer  # noqa: E402
from gsorb_slam_b200.scene import make_scene  # noqa: E402


def pose(i, n):
    a = 0.12 * np.sin(2 * np.pi * i / n)
    b = 0.06 * np.cos(2 * np.pi * i / n)
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    T = np.eye(4)
    T[:3, :3] = Ry @ Rx
    T[:3, 3] = [0.25 * np.sin(2 * np.pi * i / n), 0.1 * np.cos(2 * np.pi * i / n), 0.05 * np.sin(4 * np.pi * i / n)]
    return torch.from_numpy(T.astype(np.float32))


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 10.0 * np.log10(1.0 / max(mse, 1e-20))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=200_000)
    ap.add_argument("--G", type=int, default=8)
    ap.add_argument("--steps", type=int, default=60, help="Adam steps of the sharded schedule (the reference's Mapping.numIters)")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sc = make_scene(args.P, "tum", seed=args.seed)
    cam = sc.cam
    W, H = cam.width, cam.height
    mk = lambda m, c, o, s, q, lr=None: MapOptimizer(m, c, o, s, q, width=W, height=H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                                      projmatrix=cam.projmatrix, device=dev, lr=lr)
    gt = mk(sc.means3D, sc.colors, sc.logit_opacities, sc.log_scales, sc.unnorm_quats)
    n_train, n_held = 16, 4
    train = [pose(i, n_train).to(dev) for i in range(n_train)]
    held = [pose(i + 0.5, n_train).to(dev) for i in range(0, n_train, n_train // n_held)]
    frames = []
    for T in train + held:
        c, ds, _, _ = gt.render_fused(T)
        frames.append((c.clone().clamp(0, 1), ds[0].clone()))
    held_frames = frames[n_train:]
    rng = np.random.default_rng(args.seed + 1)
    start = dict(m=sc.means3D + rng.normal(0, 0.004, sc.means3D.shape).astype(np.float32) * sc.means3D[:, 2:3],
                 c=np.clip(sc.colors + rng.normal(0, 0.15, sc.colors.shape), 0, 1).astype(np.float32),
                 o=(sc.logit_opacities + rng.normal(0, 0.5, sc.P)).astype(np.float32),
                 s=(sc.log_scales + rng.normal(0, 0.15, sc.log_scales.shape)).astype(np.float32),
                 q=(sc.unnorm_quats + rng.normal(0, 0.05, sc.unnorm_quats.shape)).astype(np.float32))
    G, S = args.G, args.steps
    seq = rng.integers(0, n_train, G * S)

    def held_psnr(mo):
        return float(np.mean([psnr(mo.render_fused(T)[0].clamp(0, 1), f[0]) for T, f in zip(held, held_frames)]))

    def run(group, reduce, lr_mul, beta_pow=1):
        lr = {k: v * lr_mul for k, v in DEFAULT_LR.items()}
        mo = mk(start["m"], start["c"], start["o"], start["s"], start["q"], lr)
        mo.betas = (0.9 ** beta_pow, 0.999 ** beta_pow)   # beta^G: the moment averages forget over the same number of FRAMES
        p0 = held_psnr(mo)
        acc = torch.zeros_like(mo.grads.flat)
        for i in range(0, len(seq), group):
            acc.zero_()
            for k in seq[i:i + group]:
                mo.slam_gradients(train[k], frames[k][0], frames[k][1])
                acc += mo.grads.flat
            mo.grads.flat.copy_(acc if reduce == "sum" else acc / group)
            mo.adam()
        return p0, held_psnr(mo)

    out = {"P": args.P, "G": G, "adam_steps_sharded": S, "frames_consumed": int(len(seq)), "train_views": n_train, "held_out_views": n_held,
           "image": f"{W}x{H}", "loss": "MapOptimizer.slam_gradients (L1 + SSIM + depth, replica.yaml weights)", "runs": []}
    p0, ref = run(1, "sum", 1.0)
    out["psnr_start_db"] = p0
    out["reference_schedule"] = {"what": f"{len(seq)} single-view Adam steps (G = 1)", "psnr_db": ref}
    # same NUMBER of Adam steps as the shard, single view each: what one GPU reaches in the shard's wall time
    mo_lr = {k: v for k, v in DEFAULT_LR.items()}
    lr_rules = [("sum", 1.0, 1), ("mean", 1.0, 1), ("mean", float(np.sqrt(G)), 1), ("mean", 0.5 * G, 1), ("mean", float(G), 1),
                ("mean", 1.25 * G, 1), ("mean", 1.5 * G, 1), ("mean", 2.0 * G, 1),
                ("mean", 0.5 * G, G), ("mean", float(G), G), ("mean", 1.25 * G, G), ("mean", 1.5 * G, G)]
    for reduce, mul, bp in lr_rules:
        _, p = run(G, reduce, mul, bp)
        out["runs"].append({"reduce": reduce, "lr_multiplier": mul, "betas": [0.9 ** bp, 0.999 ** bp], "psnr_db": p,
                            "delta_vs_reference_db": p - ref})
    seq_short = seq[:S]
    seq_full, seq = seq, seq_short
    _, p1 = run(1, "sum", 1.0)
    seq = seq_full
    out["single_gpu_same_wall_time"] = {"what": f"{S} single-view Adam steps (what one GPU does while the shard does its {S} steps)", "psnr_db": p1}
    best = min(out["runs"], key=lambda r: abs(r["delta_vs_reference_db"]))
    out["best_rule"] = best
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
