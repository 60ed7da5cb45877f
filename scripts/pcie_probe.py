import torch, time
for mb in (8, 85, 256):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for direction in ("h2d", "d2h"):
        for _ in range(3):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True)); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{direction} {mb} MB pinned: {ms:.3f} ms  {n/ms/1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
n = 85 * 1024 * 1024
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print(f"duplex 85MB each way: {dt*1e3:.3f} ms  {n/dt/1e9:.1f} GB/s per direction")
import subprocess; print(subprocess.run("nvidia-smi -q | grep -A6 'GPU Link Info' | head -12; nvidia-smi topo -m | head -6", shell=True, capture_output=True, text=True).stdout)
