import torch, time
dev="cuda"
for mb in (1, 4, 8, 64):
    n = mb*1024*1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f"H2D pinned {mb} MiB: {n/ms/1e6:.1f} GB/s ({ms*1e3:.1f} us)")
    e0.record()
    for _ in range(20): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f"D2H pinned {mb} MiB: {n/ms/1e6:.1f} GB/s")
import subprocess
print(subprocess.run("nvidia-smi topo -m; nvidia-smi -q | grep -i -A3 'pcie' | head -40; lscpu | head -25; numactl -H 2>/dev/null | head", shell=True, capture_output=True, text=True).stdout)
