"""Raw pinned-memory copy bandwidth of the box (context for the e2e number)."""
import torch
dev = torch.device("cuda:0")
for mb in (12.6, 50.3, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        print(f"{name} {mb:6.1f} MB: {n * 10 / (s.elapsed_time(e) * 1e-3) / 1e9:.1f} GB/s")
# both directions at once
h1 = torch.empty(50_300_000, dtype=torch.uint8).pin_memory(); d1 = torch.empty_like(h1, device=dev)
h2 = torch.empty(15_800_000, dtype=torch.uint8).pin_memory(); d2 = torch.empty_like(h2, device=dev)
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.current_stream().wait_stream(s2)
e.record()
torch.cuda.synchronize()
print(f"duplex: D2H 50.3 MB + H2D 15.8 MB per iteration: {s.elapsed_time(e) / 10:.3f} ms/iter")
