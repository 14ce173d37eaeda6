import torch, time
x = torch.empty(183_000_000, dtype=torch.uint8, pin_memory=True)
d = torch.empty_like(x, device='cuda')
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10): d.copy_(x, non_blocking=True)
    e1.record(s); torch.cuda.synchronize()
print("H2D pinned GB/s", 10*183e6/ (e0.elapsed_time(e1)/1e3)/1e9)
h = torch.empty(17_000_000, dtype=torch.uint8, pin_memory=True)
with torch.cuda.stream(s):
    e0.record(s)
    for _ in range(10): h.copy_(d[:17_000_000], non_blocking=True)
    e1.record(s); torch.cuda.synchronize()
print("D2H pinned GB/s", 10*17e6/(e0.elapsed_time(e1)/1e3)/1e9)
