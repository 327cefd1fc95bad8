"""Write-only / read-only / copy bandwidth probe (context for the roofline: the assembly kernel is a pure writer)."""
import torch, time
n = 1_296_432_036
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n // 2, dtype=torch.float64, device="cuda")
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_()); print(f"memset (zero_) {n*8/1e9:.2f} GB: {ms:.3f} ms -> {n*8/ms/1e6:.0f} GB/s")
ms = t(lambda: x.fill_(1.5)); print(f"fill_ {n*8/1e9:.2f} GB: {ms:.3f} ms -> {n*8/ms/1e6:.0f} GB/s")
ms = t(lambda: y.copy_(x[: n // 2])); print(f"copy {n//2*8/1e9:.2f} GB r + w: {ms:.3f} ms -> {n//2*16/ms/1e6:.0f} GB/s")
ms = t(lambda: x.sum()); print(f"read (sum) {n*8/1e9:.2f} GB: {ms:.3f} ms -> {n*8/ms/1e6:.0f} GB/s")
