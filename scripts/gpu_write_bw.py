"""Write-only HBM bandwidth ceiling (memset / fill of a conv1-phase-image-sized buffer) next to a device copy."""
import torch
n = 326 * 1000 * 1000
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        big.zero_()  # evict
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
ms = t(lambda: a.zero_()); print("memset  %.1f us  %.0f GB/s written" % (ms * 1e3, n / ms / 1e6))
a16 = a.view(torch.bfloat16)
ms = t(lambda: a16.fill_(1.0)); print("fill    %.1f us  %.0f GB/s written" % (ms * 1e3, n / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy    %.1f us  %.0f GB/s read+written" % (ms * 1e3, 2 * n / ms / 1e6))
