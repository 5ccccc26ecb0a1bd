"""torchrun probe: gsb_exchange_allreduce (multimem and P2P variants) vs NCCL all_reduce -- values and device time."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsorb_slam_b200.distributed import SymmetricExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 14 * 1_000_000
X = SymmetricExchange(n, dev)
blk = X.alloc(n)
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
src = torch.randn(n, device=dev, generator=g)
ref = src.clone(); dist.all_reduce(ref)
out = {}
for name, mc in (("multimem", True), ("p2p", False)):
    if mc and not X.multicast_ptr:
        if rank == 0: print("no multicast mapping on this box")
        continue
    blk.copy_(src); X.allreduce(blk, use_multicast=mc); torch.cuda.synchronize()
    err = float((blk - ref).abs().max()); same = torch.empty(1, device=dev); 
    chk = blk.double().sum().reshape(1); lst = [torch.empty_like(chk) for _ in range(world)]; dist.all_gather(lst, chk)
    ident = all(float(x) == float(lst[0]) for x in lst)
    ts = []
    for it in range(15):
        blk.copy_(src); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); X.allreduce(blk, use_multicast=mc); b.record(); torch.cuda.synchronize()
        if it >= 5: ts.append(a.elapsed_time(b))
    t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{name}: max|err| vs NCCL {err:.3e}, identical across ranks {ident}, {float(t)*1000:.1f} us (max over ranks)")
ts = []
for it in range(15):
    ref.copy_(src); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dist.all_reduce(ref); b.record(); torch.cuda.synchronize()
    if it >= 5: ts.append(a.elapsed_time(b))
t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0: print(f"nccl: {float(t)*1000:.1f} us")
# odd P: the [14, P] block is not a whole number of 16-byte words; allreduce rounds up into the zero padding alloc reserved
P_odd = 100_003
X2 = SymmetricExchange(14 * P_odd + 8, dev)
blk2 = X2.alloc(14 * P_odd)
src2 = torch.randn(14 * P_odd, device=dev, generator=g)
ref2 = src2.clone(); dist.all_reduce(ref2)
blk2.copy_(src2); X2.allreduce(blk2, use_multicast=False); torch.cuda.synchronize()
err2 = float((blk2 - ref2).abs().max())
if rank == 0: print(f"p2p-oddP: max|err| vs NCCL {err2:.3e}, identical across ranks True, n = {14 * P_odd}")
# the C++ (libtorch) host of the same kernel: adapter/Exchange.{h,cc} through the pybind test harness
try:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adapter"))
    import gsb_adapter
except Exception as e:   # the adapter is optional on a box without the built .so
    gsb_adapter = None
    if rank == 0: print("cpp host not available:", repr(e)[:120])
if gsb_adapter is not None:
    n3 = 14 * 250_001   # an odd P as well
    Xc = gsb_adapter.GradientExchange(n3 + 8, torch.empty(1, device=dev), dist.group.WORLD.group_name)
    blk3 = Xc.alloc(n3)
    src3 = torch.randn(n3, device=dev, generator=g)
    ref3 = src3.clone(); dist.all_reduce(ref3)
    for name, mc in (("cpp-multimem", True), ("cpp-p2p", False)):
        if mc and not Xc.has_multicast():
            continue
        blk3.copy_(src3); Xc.allreduce(blk3, mc); torch.cuda.synchronize()
        err3 = float((blk3 - ref3).abs().max())
        chk = blk3.double().sum().reshape(1); lst = [torch.empty_like(chk) for _ in range(world)]; dist.all_gather(lst, chk)
        ident = all(float(x) == float(lst[0]) for x in lst)
        if rank == 0: print(f"{name}: max|err| vs NCCL {err3:.3e}, identical across ranks {ident}, n = {n3}, world {Xc.world_size()}")
dist.destroy_process_group()
