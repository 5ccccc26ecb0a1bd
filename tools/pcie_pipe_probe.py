"""Pipelined copy probe: per slot stream  [5 uploads 56 MB] -> [kernel ~0.6 ms] -> [downloads 65 MB], `depth` slots round-robin."""
import torch, json, sys
dev = torch.device("cuda:0")
MB = 1 << 20
def mk(depth):
    S = []
    for _ in range(depth):
        s = dict(st=torch.cuda.Stream(),
                 uh=[torch.empty(n * MB, dtype=torch.uint8).pin_memory() for n in (12, 12, 4, 12, 16, 4)],
                 dh=[torch.empty(n * MB, dtype=torch.uint8).pin_memory() for n in (4, 1, 4, 56)])
        s["ud"] = [torch.empty_like(x, device=dev) for x in s["uh"]]; s["dd"] = [torch.empty_like(x, device=dev) for x in s["dh"]]
        s["work"] = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        S.append(s)
    return S
def run(S, n, kernel=True, sync_reuse=True):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(); a.record()
    for s in S: s["st"].wait_event(a)
    for i in range(n):
        s = S[i % len(S)]
        if sync_reuse: s["st"].synchronize()
        with torch.cuda.stream(s["st"]):
            for d, h in zip(s["ud"], s["uh"]): d.copy_(h, non_blocking=True)
            if kernel:
                for _ in range(4): s["work"].zero_()      # ~4 GB of writes ~ 0.6 ms
            for h, d in zip(s["dh"], s["dd"]): h.copy_(d, non_blocking=True)
    for s in S:
        e = torch.cuda.Event(); e.record(s["st"]); cur.wait_event(e)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
out = {}
for depth in (1, 2, 3, 4):
    S = mk(depth); run(S, 4)
    out[f"depth{depth}_kernel"] = run(S, 24); out[f"depth{depth}_copies_only"] = run(S, 24, kernel=False)
    out[f"depth{depth}_kernel_nosync"] = run(S, 24, sync_reuse=False)
    del S
print(json.dumps(out, indent=1))
