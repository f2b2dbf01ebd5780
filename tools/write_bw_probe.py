"""Write-only HBM bandwidth probe (what a pure triplet-stream writer can reach): torch fill of 12 GiB."""
import torch
x = torch.empty(12 << 30, dtype=torch.uint8, device="cuda")
for _ in range(3):
    x.zero_()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(5):
    a.record(); x.zero_(); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("write-only fill: %.2f ms for 12 GiB = %.1f GB/s" % (best, (12 << 30) / best / 1e6))
y = torch.empty(6 << 30, dtype=torch.uint8, device="cuda"); z = torch.empty(6 << 30, dtype=torch.uint8, device="cuda")
best = 1e9
for _ in range(5):
    a.record(); z.copy_(y); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("copy: %.2f ms for 6+6 GiB = %.1f GB/s" % (best, (12 << 30) / best / 1e6))
