"""CPU suite, world_size 2 over gloo: the host-side logic of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsorb_slam_b200.distributed import BLOCK_ROWS, GradBlock, allreduce_gradients, broadcast_densification, shard_keyframes
    try:
        blk = GradBlock(P, "cpu")
        g = torch.Generator().manual_seed(100 + rank)
        for name in ("means", "rgb", "opacity", "scales", "quats"):
            blk[name].copy_(torch.randn(blk[name].shape, generator=g))
        mine = blk.flat.clone()
        allreduce_gradients(blk)
        # expected: sum over ranks of the per-rank blocks (regenerate the other rank's block)
        exp = torch.zeros_like(mine)
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            b2 = GradBlock(P, "cpu")
            for name in ("means", "rgb", "opacity", "scales", "quats"):
                b2[name].copy_(torch.randn(b2[name].shape, generator=gr))
            exp += b2.flat
        assert torch.allclose(blk.flat, exp, atol=1e-6), "all-reduce of the packed block != sum of rank blocks"
        # views alias the flat buffer (no pack / unpack around the collective)
        assert blk["quats"].data_ptr() == blk.flat.data_ptr() + 10 * P * 4
        assert blk.flat.numel() == BLOCK_ROWS * P
        # average mode
        blk2 = GradBlock(P, "cpu")
        blk2.flat.fill_(float(rank + 1))
        allreduce_gradients(blk2, average=True)
        assert torch.allclose(blk2.flat, torch.full_like(blk2.flat, sum(range(1, world + 1)) / world))
        # keyframe sharding: disjoint cover
        kfs = list(range(7))
        mine_kf = shard_keyframes(kfs)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine_kf)
        assert sorted(sum(gathered, [])) == kfs
        # densification broadcast: everyone ends with rank 0's rows
        rows = torch.arange(3 * BLOCK_ROWS, dtype=torch.float32).reshape(3, BLOCK_ROWS) if rank == 0 else None
        out = broadcast_densification(rows)
        assert out.shape == (3, BLOCK_ROWS) and float(out[2, 13]) == 3 * BLOCK_ROWS - 1
    finally:
        dist.destroy_process_group()


def test_world2_gloo_exchange_step():
    mp.spawn(_worker, args=(2, _free_port(), 1000), nprocs=2, join=True)


def test_tile_row_bands_cover_and_balance():
    from gsorb_slam_b200.distributed import tile_row_bands
    for tiles_y, world in [(61, 2), (61, 4), (61, 8), (30, 8), (3, 8), (43, 1)]:
        b = tile_row_bands(tiles_y, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == tiles_y
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(e >= s for s, e in b)
        if tiles_y >= world:
            h = [e - s for s, e in b]
            assert max(h) - min(h) <= 2
    # load-balanced: a heavy first row gets a band of its own
    b = tile_row_bands(8, 2, weights=[10, 1, 1, 1, 1, 1, 1, 1])
    assert b == [(0, 1), (1, 8)]
    with pytest.raises(ValueError):
        tile_row_bands(4, 0)


def test_pose_chain_rule_and_reprojection_gradient_match_torch_autograd():
    """The host-side closed forms of gsorb_slam_b200/tracking.py (Rt2T and its backward, the ORB reprojection term and its
    gradient w.r.t. Tcw; src/Utils.cc:170-179, src/Render.cc:1058-1065) against torch autograd of the literal expressions."""
    import numpy as np
    import torch
    from gsorb_slam_b200.tracking import PoseOptimizer, rt2T, rt2T_backward_np, rt2T_np
    rng = np.random.default_rng(0)
    q, t, G = rng.normal(size=4).astype(np.float32), rng.normal(size=3).astype(np.float32), rng.normal(size=(3, 4)).astype(np.float32)
    qt, tt = torch.tensor(q, requires_grad=True), torch.tensor(t, requires_grad=True)
    T = rt2T(qt, tt)
    (T[:3] * torch.tensor(G)).sum().backward()
    gq, gt = rt2T_backward_np(q, G)
    assert np.abs(rt2T_np(q, t) - T.detach().numpy()).max() <= 1e-6
    assert np.abs(gq - qt.grad.numpy()).max() <= 1e-5 and np.abs(gt - tt.grad.numpy()).max() <= 1e-6
    po = PoseOptimizer.__new__(PoseOptimizer)   # the reprojection term needs no rasterizer
    M = 64
    K = np.array([[260, 0, 159.5], [0, 258, 119.5], [0, 0, 1]], np.float32)
    Xw = np.stack([rng.uniform(-1, 1, M), rng.uniform(-1, 1, M), rng.uniform(2, 5, M)], 1).astype(np.float32)
    obs, w = rng.uniform(0, 300, (M, 2)).astype(np.float32), rng.uniform(0.3, 1, M).astype(np.float32)
    po.features = dict(K=K, Xw=Xw, obs=obs, w=w, inlier=rng.random(M) < 0.8)
    Tn = rt2T_np(np.array([1, 0.01, -0.02, 0.03], np.float32), np.array([0.1, -0.05, 0.02], np.float32))
    err, Gf = po._reprojection_np(Tn, with_grad=True)
    Tt = torch.tensor(Tn, requires_grad=True)
    Xc = torch.tensor(Xw) @ Tt[:3, :3].T + Tt[:3, 3]
    e = ((Xc / Xc[:, 2:3]) @ torch.tensor(K).T)[:, :2] - torch.tensor(obs)
    er = (e * e).sum(1) * torch.tensor(w)
    er[torch.tensor(po.features["inlier"])].sum().backward()
    assert np.abs(err - er.detach().numpy()).max() <= 1e-6 * np.abs(err).max()
    assert np.abs(Gf - Tt.grad.numpy()[:3]).max() <= 1e-5 * np.abs(Gf).max()


@pytest.mark.parametrize("G", [1, 2, 8])
def test_map_keyframes_draws_one_keyframe_per_rank_and_iteration(monkeypatch, G):
    """MapOptimizer.map_keyframes (the loop of Render::RenderForFrame, src/Render.cc:420-424): every iteration draws G keyframes
    from ONE generator seeded alike on every rank and rank r takes the r-th, so the ranks of a step see different views of
    the same window and a single rank reproduces the reference's one uniform draw per iteration.  Host logic only: the
    iteration itself is replaced by a recorder."""
    import random
    import types
    from gsorb_slam_b200 import mapping
    window = [(torch.eye(4) * (k + 1), torch.full((3, 2, 2), float(k)), torch.full((2, 2), float(k))) for k in range(5)]
    iters, seed = 12, 7
    want = random.Random(seed)
    draws = [[want.randrange(len(window)) for _ in range(G)] for _ in range(iters)]
    for rank in range(G):
        monkeypatch.setattr(mapping, "world", lambda rank=rank: (rank, G))
        seen = []

        def step_slam(T, c, z, average=False, **kw):
            seen.append((int(T[0, 0]) - 1, average, kw))
            return torch.zeros(8)
        mo = types.SimpleNamespace(dev=torch.device("cpu"), step_slam=step_slam)
        terms = mapping.MapOptimizer.map_keyframes(mo, window, iters=iters, rng=random.Random(seed), w_depth=0.5)
        assert terms.shape == (8,)
        assert [s[0] for s in seen] == [d[rank] for d in draws]
        assert all(s[1] == (G > 1) and s[2] == {"w_depth": 0.5} for s in seen)   # gradients averaged over the ranks of a step
    with pytest.raises(ValueError):
        mapping.MapOptimizer.map_keyframes(types.SimpleNamespace(dev=torch.device("cpu"), step_slam=None), [], iters=1)
