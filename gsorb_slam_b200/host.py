"""Host-buffer frames through the C ABI, several in flight: the streaming front end of ``gsb_forward_backward_host_async``.

A GSORB-SLAM process that keeps its Gaussians in host memory (or a test harness, or the bench's ``e2e`` leg) pays two PCIe
transfers per frame: ~56 B per Gaussian up, the same down.  One frame alone is a serial chain upload -> kernels -> download;
with ``depth`` frames in flight -- each on its own stream, with its own device scratch and its own pinned output buffers -- the
upload of frame i+1 and the kernels of frame i+2 run under the download of frame i (PCIe is full duplex, the copy engines are
separate from the SMs), so the steady state costs max(upload, kernels, download) per frame instead of their sum.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from ._lib import GradOutputs, RasterArgs


class HostSlot:
    """One frame in flight: stream, device scratch, pinned outputs (image, depth, radii, packed [14, P] gradient block)."""

    def __init__(self, P: int, W: int, H: int, max_rendered: int, device, scratch_bytes: int):
        self.stream = torch.cuda.Stream(device=device)
        self.scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=device)
        self.color = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
        self.depth = torch.empty((1, H, W), dtype=torch.float32).pin_memory()
        self.radii = torch.empty(P, dtype=torch.int32).pin_memory()
        self.block = torch.empty(14 * P, dtype=torch.float32).pin_memory()   # means3D 3 | colour 3 | opacity 1 | scale 3 | rotation 4
        self.status = torch.zeros(4, dtype=torch.int32).pin_memory()         # num_rendered, binned, overflow latch
        g = GradOutputs()
        b = self.block.data_ptr()
        g.dL_dmean3D, g.dL_dcolor, g.dL_dopacity, g.dL_dscale, g.dL_drot = b, b + 12 * P, b + 24 * P, b + 28 * P, b + 40 * P
        self.grads = g
        self.busy = False


class HostPipeline:
    """``submit`` enqueues one forward + backward over HOST arrays and returns immediately; ``wait`` hands back the oldest
    frame's slot once its downloads have landed.  At most ``depth`` frames are in flight."""

    def __init__(self, P: int, width: int, height: int, max_rendered: Optional[int] = None, depth: int = 3, device="cuda:0"):
        self.L = _lib.lib()
        self.dev = torch.device(device)
        self.P, self.W, self.H = int(P), int(width), int(height)
        self.max_rendered = int(max_rendered) if max_rendered else 4 * self.P + 4096
        self.nscratch = int(self.L.gsb_host_scratch_bytes(self.P, 0, self.W, self.H, self.max_rendered))
        self.slots: List[HostSlot] = [HostSlot(self.P, self.W, self.H, self.max_rendered, self.dev, self.nscratch) for _ in range(depth)]
        self._next = 0
        self._queue: List[HostSlot] = []

    def submit(self, host_args: RasterArgs, dL_dpix_host_ptr: int) -> HostSlot:
        """``host_args``: a ``gsb_raster_args`` whose pointers are HOST (pinned) arrays, kept alive by the caller."""
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        if slot.busy:
            self.wait()   # the ring is full: retire the oldest frame (it is this slot)
        with torch.cuda.device(self.dev):
            _lib.check(self.L.gsb_forward_backward_host_async(
                C.byref(host_args), self.max_rendered, dL_dpix_host_ptr, slot.color.data_ptr(), slot.depth.data_ptr(),
                slot.radii.data_ptr(), C.byref(slot.grads), slot.scratch.data_ptr(), self.nscratch, slot.status.data_ptr(),
                slot.stream.cuda_stream))
        slot.busy = True
        self._queue.append(slot)
        return slot

    def wait(self) -> Optional[HostSlot]:
        """Oldest frame in flight: blocks until its stream has drained; raises if the binning capacity overflowed."""
        if not self._queue:
            return None
        slot = self._queue.pop(0)
        slot.stream.synchronize()
        slot.busy = False
        if int(slot.status[2]) != 0:
            raise ValueError(f"GSB_ERR_OVERFLOW: num_rendered {int(slot.status[0]) & 0xffffffff} exceeded the binning capacity {self.max_rendered}")
        return slot

    def drain(self) -> None:
        while self._queue:
            self.wait()
