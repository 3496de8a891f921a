"""Read-only / write-only / copy HBM streams (torch kernels, CUDA events, 4 GB buffers >> L2): what a pure write stream can
reach on this B200, next to the copy peak that MEASURED_PEAKS.json holds."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 30   # fp32 elements = 4 GiB
a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)
def t(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
G = 4 * n / 1e6
ms = t(lambda: a.zero_()); print("memset (zero_)      %.3f ms  %.0f GB/s written" % (ms, G / ms))
ms = t(lambda: a.fill_(1.5)); print("fill_ kernel        %.3f ms  %.0f GB/s written" % (ms, G / ms))
ms = t(lambda: b.copy_(a)); print("copy                %.3f ms  %.0f GB/s read+written" % (ms, 2 * G / ms))
ms = t(lambda: a.sum()); print("sum (read only)     %.3f ms  %.0f GB/s read" % (ms, G / ms))
ms = t(lambda: torch.add(a, 1.0, out=b)); print("add scalar (1r+1w)  %.3f ms  %.0f GB/s read+written" % (ms, 2 * G / ms))
ms = t(lambda: torch.add(a, b, out=b)); print("add (2r+1w)         %.3f ms  %.0f GB/s read+written" % (ms, 3 * G / ms))
