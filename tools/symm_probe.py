"""Probe: does torch symmetric memory (CUDA VMM peer mapping over NVLink) work on this box?  torchrun, 2+ ranks."""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 16 << 20
t = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "signal", len(hdl.signal_pad_ptrs), flush=True)
t.fill_(rank + 1.0)
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
print(rank, "peer value", float(peer[123].item()), flush=True)
# peer write bandwidth with a plain copy kernel
src = torch.ones(n, device=dev)
torch.cuda.synchronize(); hdl.barrier(channel=0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): peer.copy_(src)
a.record()
for _ in range(10): peer.copy_(src)
b.record(); torch.cuda.synchronize()
print(rank, "peer copy GB/s", n * 4 * 10 / (a.elapsed_time(b) * 1e-3) / 1e9, flush=True)
hdl.barrier(channel=0)
dist.destroy_process_group()
