"""What the host link of this box sustains: H2D alone, D2H alone, both at once (pinned memory, 60 MB / 65 MB per direction and step)."""
import torch, time, json
dev = torch.device("cuda:0")
up_h = torch.empty(60 << 20, dtype=torch.uint8).pin_memory(); up_d = torch.empty(60 << 20, dtype=torch.uint8, device=dev)
dn_h = torch.empty(65 << 20, dtype=torch.uint8).pin_memory(); dn_d = torch.empty(65 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, n=20):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); s1.wait_event(a); s2.wait_event(a)
    for _ in range(n):
        if up:
            with torch.cuda.stream(s1): up_d.copy_(up_h, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2): dn_h.copy_(dn_d, non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event(); e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
run(True, True, 3)
out = {"h2d_only_ms": run(True, False), "d2h_only_ms": run(False, True), "both_ms": run(True, True)}
out["h2d_GBps"] = 60 * 1.048576e-3 / out["h2d_only_ms"] * 1e0; out["d2h_GBps"] = 65 * 1.048576e-3 / out["d2h_only_ms"]
out["both_total_GBps"] = 125 * 1.048576e-3 / out["both_ms"]
print(json.dumps(out))
