"""Pinned host -> device copy bandwidth of the box: one stream vs the same bytes split over 2 / 4 streams."""
import torch

dev = torch.device("cuda:0")
n = 49_235_968 // 4
h = torch.zeros(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
for parts in (1, 2, 4):
    step = (n + parts - 1) // parts
    for rep in range(3):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for it in range(10):
            for p in range(parts):
                st = streams[p]
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    d[p * step:(p + 1) * step].copy_(h[p * step:(p + 1) * step], non_blocking=True)
            for p in range(parts):
                torch.cuda.current_stream().wait_stream(streams[p])
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
    print(f"{parts} stream(s): {ms:.3f} ms per {n * 4 / 1e6:.1f} MB = {n * 4 / ms / 1e6:.1f} GB/s")
# device -> host for reference
hh = torch.empty(n, dtype=torch.float32).pin_memory()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for it in range(10):
    hh.copy_(d, non_blocking=True)
e.record()
torch.cuda.synchronize()
print(f"D2H: {n * 4 / (s.elapsed_time(e) / 10) / 1e6:.1f} GB/s")
