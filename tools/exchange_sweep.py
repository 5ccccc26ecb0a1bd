"""torchrun probe (tune build): gsb_exchange_allreduce over CTAs-per-SM x unroll, multimem and P2P, 56 MB block."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
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
def timed(fn):
    ts = []
    for it in range(12):
        blk.copy_(src); dist.barrier(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if it >= 4: ts.append(a.elapsed_time(b))
    t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1000
for mc in ((True, False) if X.multicast_ptr else (False,)):
    for c in (1, 2, 4, 6):
        for u in (2, 4, 8):
            os.environ["GSB_XCH_CTAS_PER_SM"] = str(c); os.environ["GSB_XCH_UNROLL"] = str(u)
            us = timed(lambda: X.allreduce(blk, use_multicast=mc))
            err = float((blk - ref).abs().max())
            if rank == 0: print(f"{'multimem' if mc else 'p2p'} ctas/sm {c} unroll {u}: {us:.1f} us, max|err| {err:.2e}", flush=True)
# fixed cost of a launch: the two cross-rank handshakes + fence, on a 16 KB block
small = X.buf[:4096]
for mc in ((True, False) if X.multicast_ptr else (False,)):
    os.environ["GSB_XCH_CTAS_PER_SM"] = "2"; os.environ["GSB_XCH_UNROLL"] = "4"
    us = timed(lambda: X.allreduce(small, use_multicast=mc))
    if rank == 0: print(f"{'multimem' if mc else 'p2p'} 16 KB block (handshakes + fence only): {us:.1f} us", flush=True)
blk_n = src.clone()
us = timed(lambda: dist.all_reduce(blk))
if rank == 0: print(f"nccl: {us:.1f} us")
dist.destroy_process_group()
