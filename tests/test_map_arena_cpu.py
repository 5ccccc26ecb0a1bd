"""CPU suite: the host logic of the growable map (gsorb_slam_b200/mapping.py, distributed.GradBlock) -- the capacity-doubling
arenas that replace Gaussian::CatTensorToOptimizer's torch::cat per tensor and per Adam moment (src/Gaussian.cc:50-95, 241-258).
No kernel is launched: MapOptimizer is built on the CPU device, where only its tensor bookkeeping runs."""
import numpy as np
import pytest
import torch

from gsorb_slam_b200.distributed import BLOCK_ROWS, GROUPS, GradBlock
from gsorb_slam_b200.mapping import MapOptimizer, keyframe_batch_hyperparameters


def _rows(P, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return r(P, 3), r(P, 3), r(P), r(P, 3), r(P, 4)


def _mk(P, **kw):
    return MapOptimizer(*_rows(P, 1), width=64, height=48, tanfovx=0.5, tanfovy=0.4, projmatrix=np.eye(4, dtype=np.float32),
                        device="cpu", **kw)


def test_grad_block_arena_layout_and_growth():
    b = GradBlock(5, "cpu", capacity=8)
    assert b.flat.numel() == BLOCK_ROWS * 8 and b.group_sizes() == [w * 8 for _, w in GROUPS]
    off = 0
    for name, w in GROUPS:   # group g starts at offset_g * capacity floats; only P rows are in use
        assert b[name].shape == (5, w) and b.ptr(name) == b.flat.data_ptr() + 4 * off * 8
        b[name].fill_(off + 1.0)
        off += w
    assert float(b.flat.sum()) == sum((o + 1.0) * 5 * w for o, (_, w) in zip((0, 3, 6, 7, 10), GROUPS))   # padding rows stay zero
    b.resize(7)
    assert b["quats"].shape == (7, 4) and float(b["quats"][5:].abs().max()) == 0.0 and float(b["quats"][:5].min()) == 11.0
    g = b.grown(16)
    assert g.capacity == 16 and g.P == 7 and torch.equal(g["rgb"], b["rgb"]) and float(g.flat.sum()) == float(b.flat.sum())
    with pytest.raises(ValueError):
        b.resize(9)
    with pytest.raises(ValueError):
        GradBlock(9, "cpu", capacity=8)


@pytest.mark.parametrize("P", [1, 6, 1001])
def test_automatic_capacity_keeps_every_group_16_byte_aligned(P):
    """The fused map update stages rows with bulk copies only when every group of every arena starts 16-byte aligned
    (csrc/map_update.cu): an automatic capacity is a multiple of 4 rows, an explicit one is taken as given."""
    mo = _mk(P)
    assert mo.capacity % 4 == 0 and mo.capacity >= P
    for blk in (mo.params, mo.grads, mo.exp_avg, mo.exp_avg_sq):
        assert all((blk.ptr(n) - blk.flat.data_ptr()) % 16 == 0 for n, _ in GROUPS)
    assert _mk(P, capacity=P + 1).capacity == P + 1


def test_add_gaussians_appends_like_cat_with_zero_moments():
    P, K = 10, 7
    mo = _mk(P)                                        # capacity 12: the append must grow the arenas
    before = {n: mo.params[n].clone() for n, _ in GROUPS}
    for n, _ in GROUPS:                                # pretend some optimisation happened
        mo.exp_avg[n].fill_(0.5)
        mo.exp_avg_sq[n].fill_(0.25)
    mo.t = 3
    new = _rows(K, 2)
    assert mo.add_gaussians(*new) == K
    assert mo.P == P + K and mo.capacity >= P + K and mo.capacity % 4 == 0
    assert mo.capacity >= 12 + 12 // 2                 # grows by half at least (amortised appends)
    want = dict(means=new[0], rgb=new[1], opacity=new[2].reshape(K, 1), scales=new[3], quats=new[4])
    for n, _ in GROUPS:
        assert torch.equal(mo.params[n][:P], before[n]) and torch.equal(mo.params[n][P:], want[n])
        assert float(mo.exp_avg[n][:P].min()) == 0.5 and float(mo.exp_avg[n][P:].abs().max()) == 0.0     # CatTensorToOptimizer: zeros
        assert float(mo.exp_avg_sq[n][:P].min()) == 0.25 and float(mo.exp_avg_sq[n][P:].abs().max()) == 0.0
        assert mo.grads[n].shape == mo.params[n].shape
    assert mo.t == 3                                   # ONE step counter, shared by old and new rows
    # the per-row temporaries and the C-ABI argument block follow the arena
    assert mo.args.P == P + K and mo.means_cam.shape[0] == mo.capacity and mo.radii.shape[0] == mo.capacity
    assert mo.max_rendered >= 4 * mo.P + 4096
    # a second append that fits does not reallocate
    cap, ptr = mo.capacity, mo.params.flat.data_ptr()
    room = cap - mo.P
    if room:
        mo.add_gaussians(*_rows(room, 3))
        assert mo.capacity == cap and mo.params.flat.data_ptr() == ptr and mo.P == cap
    assert mo.add_gaussians(*_rows(0, 4)) == 0


def test_keyframe_batch_hyperparameters_rule():
    """lr x G / 2, betas ** G for a G-view minibatch step (tools/minibatch_parity.py); G = 1 is the reference's setting."""
    lr1, b1 = keyframe_batch_hyperparameters(1)
    assert lr1["means"] == 1e-4 and lr1["opacity"] == 0.05 and b1 == (0.9, 0.999)
    lr8, b8 = keyframe_batch_hyperparameters(8)
    assert lr8["scales"] == pytest.approx(4e-3) and b8[0] == pytest.approx(0.9 ** 8) and b8[1] == pytest.approx(0.999 ** 8)
