"""Multi-GPU plumbing of the render-optimise loop: one process per GPU,
``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the CPU tests).

The reference is single-GPU (every call site passes device 0, src/Render.cc:775,824).  The loop
shards two ways, both with ONE exchange step -- a sum all-reduce of the packed per-Gaussian
gradient block -- placed between ``loss.backward()`` and ``Gaussian::StepUpdataForGaussian``
(src/Render.cc:471-475):

* keyframe-batch shard: rank r renders keyframes ``keyframes[r::world]`` of the candidate list
  (``Render::RenderForFrame`` draws one random keyframe per Adam step, src/Render.cc:423) over
  replicated Gaussians;
* tile-row shard: rank r owns a contiguous band of tile rows of the same keyframe
  (``tile_row_bands``).

The gradient block is ``[14, P]`` fp32, group-major so every group is a contiguous view that
``gsb_backward`` / ``gsb_prologue_backward`` write in place and ``gsb_adam_step`` reads in
place -- no pack / unpack copy around the collective:

    rows 0-2  d(mean)   rows 3-5  d(rgb)   row 6  d(logit opacity)   rows 7-9  d(log scale)
    rows 10-13  d(unnormalised quaternion)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

GROUPS: Tuple[Tuple[str, int], ...] = (("means", 3), ("rgb", 3), ("opacity", 1), ("scales", 3), ("quats", 4))
BLOCK_ROWS = sum(w for _, w in GROUPS)  # 14


class GradBlock:
    """One contiguous fp32 buffer holding the five parameter groups, group-major, with a ``[P, w]`` view per group.

    ``capacity`` (rows, >= P) makes it an ARENA: group g starts at ``offset_g * capacity`` and only its first ``P`` rows are
    in use, so Gaussians can be appended in place (``resize``) until the capacity is exhausted -- the capacity-doubling
    replacement of the ``torch::cat`` per tensor and per Adam moment of ``Gaussian::CatTensorToOptimizer``
    (src/Gaussian.cc:241-258).  With ``capacity == P`` (default) the layout is the packed ``[14, P]`` block.  ``flat`` always
    spans the whole arena: the padding rows are zero in every block (parameters, gradients, moments), so Adam and the
    all-reduce can run over ``flat`` unchanged.  ``storage`` lets the caller place it in a symmetric allocation
    (``SymmetricExchange.alloc``)."""

    def __init__(self, P: int, device, dtype=torch.float32, storage: Optional[torch.Tensor] = None, capacity: Optional[int] = None):
        self.P = int(P)
        self.capacity = int(capacity) if capacity is not None else self.P
        if self.capacity < self.P:
            raise ValueError("capacity smaller than P")
        n = BLOCK_ROWS * self.capacity
        if storage is None:
            storage = torch.zeros(n, dtype=dtype, device=device)
        if storage.numel() < n or storage.dtype != dtype:
            raise ValueError("storage too small for the gradient block")
        self.flat = storage[:n]
        self._make_views()

    def _make_views(self):
        self.views: Dict[str, torch.Tensor] = {}
        off = 0
        for name, w in GROUPS:
            self.views[name] = self.flat[off * self.capacity:off * self.capacity + w * self.P].view(self.P, w)
            off += w

    def resize(self, P: int) -> None:
        """Change the number of rows in use (<= capacity); the views are rebuilt, the data stays where it is."""
        if P > self.capacity or P < 0:
            raise ValueError("resize beyond the arena's capacity")
        self.P = int(P)
        self._make_views()

    def grown(self, capacity: int) -> "GradBlock":
        """A new arena of ``capacity`` rows holding this one's rows (padding zero)."""
        nb = GradBlock(self.P, self.flat.device, self.flat.dtype, capacity=capacity)
        for name, _ in GROUPS:
            nb.views[name].copy_(self.views[name])
        return nb

    def group_sizes(self):
        """Elements per group over the WHOLE arena (what gsb_adam_step_groups walks)."""
        return [w * self.capacity for _, w in GROUPS]

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]

    def ptr(self, name: str) -> int:
        return self.views[name].data_ptr()


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class SymmetricExchange:
    """The exchange step as ONE libgsb kernel over NVLink / NVSwitch peer memory
    (``gsb_exchange_allreduce``, csrc/exchange.cu) instead of an NCCL call.

    ``alloc(n)`` returns an fp32 tensor inside a symmetric allocation (same offset on every rank,
    peer-mapped by ``torch.distributed._symmetric_memory``; the NVSwitch multicast mapping is used
    when the box offers one); ``allreduce(t)`` sums it in place across the ranks on the current
    stream.  All ranks must create the object and call its methods in the same order.
    Raises if symmetric memory is unavailable -- callers fall back to ``allreduce_gradients``."""

    def __init__(self, capacity_floats: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._C, self._lib, self.L = C, _lib, _lib.lib()
        self.rank, self.world = world()
        if self.world < 2:
            raise RuntimeError("SymmetricExchange needs an initialised process group with world size >= 2")
        self.group = group if group is not None else dist.group.WORLD
        self.dev = torch.device(device)
        self.capacity = (int(capacity_floats) + 3) // 4 * 4
        sync_words = int(self.L.gsb_exchange_sync_bytes(self.world)) // 4
        # one symmetric allocation: [data | handshake scratch]
        self.buf = symm_mem.empty(self.capacity + sync_words, dtype=torch.float32, device=self.dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group=self.group)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self._peer = (C.c_void_p * self.world)(*base)
        self._sync = (C.c_void_p * self.world)(*[p + 4 * self.capacity for p in base])
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.multicast_ptr = mc
        self._used = 0
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)   # every rank's scratch is zeroed before anyone handshakes

    def alloc(self, n: int) -> torch.Tensor:
        n4 = (int(n) + 3) // 4 * 4
        if self._used + n4 > self.capacity:
            raise RuntimeError("symmetric allocation exhausted")
        t = self.buf[self._used:self._used + int(n)]
        self._used += n4
        return t

    def allreduce(self, t: torch.Tensor, use_multicast: bool = True):
        """In-place sum over the ranks of a tensor obtained from ``alloc``.  The kernel moves 16-byte words: a length that
        is not a multiple of 4 floats (the ``[14, P]`` block of an odd P) is rounded up into the padding ``alloc`` reserved
        behind the tensor, which is zero on every rank and stays zero."""
        C = self._C
        off = t.data_ptr() - self.buf.data_ptr()
        n4 = (t.numel() + 3) // 4 * 4
        if off < 0 or off % 16 or off + n4 * 4 > self.capacity * 4:
            raise ValueError("tensor is not a 16-byte-aligned slice of the symmetric allocation")
        peer = (C.c_void_p * self.world)(*[int(p) + off for p in self.hdl.buffer_ptrs])
        mc = self.multicast_ptr + off if (self.multicast_ptr and use_multicast) else None
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        self._lib.check(self.L.gsb_exchange_allreduce(mc, peer, self._sync, n4, self.rank, self.world, stream))


def allreduce_gradients(block: GradBlock, average: bool = False, group=None, async_op: bool = False):
    """The exchange step: ONE sum all-reduce over the whole block (56 B per Gaussian).  With
    ``average`` the sum is divided by the world size (mini-batch mean over keyframes)."""
    rank, n = world()
    if n == 1:
        return None
    work = dist.all_reduce(block.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if average and not async_op:
        block.flat.mul_(1.0 / n)
    return work


def shard_keyframes(keyframes: List, rank: Optional[int] = None, world_size: Optional[int] = None) -> List:
    """Round-robin assignment of a candidate keyframe list to ranks (rank r takes r, r+n, ...)."""
    r, n = world()
    rank = r if rank is None else rank
    world_size = n if world_size is None else world_size
    return list(keyframes[rank::world_size])


def tile_row_bands(tiles_y: int, world_size: int, weights: Optional[List[float]] = None) -> List[Tuple[int, int]]:
    """Contiguous bands ``[begin, end)`` of tile rows, one per rank, covering ``[0, tiles_y)``.
    ``weights`` (per tile row, e.g. instance counts of the previous frame) balances the bands by
    load instead of by height; bands may be empty when there are more ranks than rows (or one row carries most of the load).
    An empty band is returned as ``(k, k)``; hand it to the rasterizer through ``band_for_rasterizer`` -- in
    ``gsb_raster_args`` the pair ``(0, 0)`` means "the whole image"."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if weights is None:
        weights = [1.0] * tiles_y
    if len(weights) != tiles_y:
        raise ValueError("one weight per tile row")
    total = float(sum(weights))
    cum = [0.0]
    for w in weights:
        cum.append(cum[-1] + float(w))
    bands, begin = [], 0
    for r in range(world_size):
        if r == world_size - 1:
            end = tiles_y
        else:
            target = total * (r + 1) / world_size   # cut where the cumulative load is closest to r+1 equal shares
            end = min(range(begin, tiles_y + 1), key=lambda e: (abs(cum[e] - target), e))
        bands.append((begin, end))
        begin = end
    return bands


def band_for_rasterizer(band: Tuple[int, int], tiles_y: int) -> Tuple[int, int]:
    """``gsb_raster_args::tile_row_begin / tile_row_end`` for a band of ``tile_row_bands``.  ``(0, 0)`` is the C ABI's
    "no shard" value (a zero-initialised args struct renders the whole image), so an EMPTY band -- which ``tile_row_bands``
    legitimately returns as ``(0, 0)`` when the first rows are heavy -- is moved to the equivalent ``(tiles_y, tiles_y)``."""
    b0, b1 = int(band[0]), int(band[1])
    if b1 <= b0:
        return (int(tiles_y), int(tiles_y))
    return (b0, b1)


def broadcast_densification(new_rows: Optional[torch.Tensor], src: int = 0, group=None) -> torch.Tensor:
    """Densify / prune decisions must be identical on every rank (replicated Gaussians): rank ``src``
    decides, everyone else receives.  ``new_rows`` is a ``[K, 14]`` tensor of raw parameters on ``src``."""
    rank, n = world()
    if n == 1:
        return new_rows
    dev = new_rows.device if new_rows is not None else torch.device("cpu")
    count = torch.tensor([0 if new_rows is None else new_rows.shape[0]], dtype=torch.int64, device=dev)
    dist.broadcast(count, src=src, group=group)
    k = int(count.item())
    buf = new_rows if rank == src else torch.empty((k, BLOCK_ROWS), dtype=torch.float32, device=dev)
    if k:
        dist.broadcast(buf, src=src, group=group)
    return buf
