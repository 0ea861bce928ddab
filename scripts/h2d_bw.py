import torch, time
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
a.fill_(1)
b = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 32 << 20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        e0.record()
        for o in range(0, n, chunk):
            b[o:o + chunk].copy_(a[o:o + chunk], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"H2D pinned chunk={chunk >> 20} MiB: {n / e0.elapsed_time(e1) * 1e-6:.1f} GB/s")
p = torch.empty(n, dtype=torch.uint8)
p.fill_(2)
t = time.perf_counter(); b.copy_(p); torch.cuda.synchronize(); print(f"H2D pageable: {n / (time.perf_counter() - t) * 1e-9:.1f} GB/s")
t = time.perf_counter(); a.copy_(b); torch.cuda.synchronize(); print(f"D2H pinned: {n / (time.perf_counter() - t) * 1e-9:.1f} GB/s")
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run("nvidia-smi topo -m | head -5; lscpu | grep -i numa", shell=True, capture_output=True, text=True).stdout)
