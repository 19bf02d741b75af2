"""Raw pinned host->device / device->host bandwidth on this box (context for the e2e number)."""
import torch
x = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(x, non_blocking=True)), ("D2H", lambda: x.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    print(name, "GB/s", 20 * 64 * 2 ** 20 / 1e9 / (a.elapsed_time(b) / 1e3))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); e = torch.empty_like(d)
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(20):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
b.record(); torch.cuda.synchronize()
print("H2D+D2H concurrent, GB/s each", 20 * 64 * 2 ** 20 / 1e9 / (a.elapsed_time(b) / 1e3))
