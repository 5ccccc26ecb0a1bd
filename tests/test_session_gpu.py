"""-m gpu: the loops around the mapping / tracking iteration (Render::InitWorld, RenderStartTraking, AddGaussian, RenderForFrame,
RemoveGaussian; src/Render.cc:402-616, 985-1141) run end to end on a synthetic RGB-D sequence (examples/slam_session.py)."""
import math
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def test_session_from_init_world_to_the_saved_map(tmp_path):
    import slam_session
    res = slam_session.run(128, 96, frames=4, init_iters=40, track_iters=40, map_iters=20, ply_path=str(tmp_path / "GaussianModel.ply"))
    # InitWorld: one Gaussian per pixel with a valid depth (Render::InitGaussianPoint, src/Render.cc:666-707), fitted to the frame
    assert res["init_gaussians"] == res["init_valid_pixels"] > 0.5 * 128 * 96
    first, last = res["init_loss_first_last"]
    assert math.isfinite(first) and math.isfinite(last) and last < first
    assert res["scene_radius"] > 0
    # every later frame: tracked, densified, mapped, pruned -- all finite, the map never shrinks below what the frames need
    assert len(res["per_frame"]) == 3
    for f in res["per_frame"]:
        for k in ("pose_err_after", "tracking_loss", "mapping_loss", "psnr_db"):
            assert math.isfinite(f[k]), (f["frame"], k)
        assert f["added"] >= 0 and f["removed"] >= 0 and f["gaussians"] > 0
    assert res["final_gaussians"] == res["per_frame"][-1]["gaussians"]
    # GaussianModel.ply (src/Utils.cc:182-280) carries the raw parameters: the reloaded map renders the same image
    assert res["ply_round_trip_max_abs_diff"] <= 1e-5


def test_init_world_runs_its_iterations_and_sets_the_scene_radius():
    """MapOptimizer.init_world (Render::InitWorld, src/Render.cc:496-553): back-projection of every valid pixel at a
    non-identity pose, `iters` Adam steps without regularisers, scene radius = max depth / Mapping.raduisDepthRatio (:705)."""
    import numpy as np
    import torch
    from gsorb_slam_b200.mapping import MapOptimizer
    from gsorb_slam_b200.scene import Camera
    from gsorb_slam_b200.tracking import rt2T_np
    dev = torch.device("cuda:0")
    W, H = 96, 64
    cam = Camera(W, H, 80.0, 78.0)
    g = torch.Generator().manual_seed(3)
    depth = (torch.rand(H, W, generator=g) * 2 + 1.5)
    depth[::5, ::3] = 0.0
    color = torch.rand(3, H, W, generator=g)
    T = torch.from_numpy(rt2T_np(np.array([0.99, 0.02, -0.03, 0.01], np.float32), np.array([0.1, -0.05, 0.2], np.float32)))
    mo = MapOptimizer.init_world(T, color, depth, cam.fx, cam.fy, (W - 1) / 2.0, (H - 1) / 2.0, iters=3, device=dev, width=W, height=H,
                                 tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, projmatrix=cam.projmatrix)
    assert mo.P == int((depth > 0).sum()) and mo.t == 3
    assert abs(mo.scene_radius - float(depth.max()) / 3.0) < 1e-6
    assert bool(torch.isfinite(mo.params.flat).all())
    # the rows are the back-projected pixels in raster order: the first valid pixel of the image is row 0
    i, j = [int(x) for x in (depth > 0).nonzero()[0]]
    z = float(depth[i, j])
    Twc = np.linalg.inv(T.numpy().astype(np.float64))
    want = Twc[:3, :3] @ np.array([(j - (W - 1) / 2.0) * z / cam.fx, (i - (H - 1) / 2.0) * z / cam.fy, z]) + Twc[:3, 3]
    got = mo.params["means"][0].cpu().numpy()
    assert np.abs(got - want).max() < 5e-3   # three Adam steps at lr 1e-4 moved it by at most 3e-4
