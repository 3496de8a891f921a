"""begin_conv / end_conv of the score net at cfg-2 size (28 images, 256x256, 128 channels): CUDA-event times and GB/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from inverseproblemwithdiffusionmodel_b200 import _lib as L
dev = torch.device("cuda", 0)
N, H, W, C = 28, 256, 256, 128
lib = L.lib()
x = torch.rand(N, H, W, device=dev)
w1 = torch.randn(C, 9, device=dev) * 0.3
b1 = torch.randn(C, device=dev)
out = torch.empty(N, H, W, C, device=dev)
stats = torch.zeros(N, C, 2, dtype=torch.float64, device=dev)
a16 = torch.randn(N, H, W, C, device=dev).half()
w9 = torch.randn(9, C, device=dev) * 0.1
bias = torch.randn(1, device=dev)
sig = torch.rand(10, device=dev) + 0.5
lab = torch.zeros(N, dtype=torch.long, device=dev)
o1 = torch.empty(N, H, W, device=dev)
ws = torch.empty(N * H * W * 9, device=dev)
st = L.stream()
first = lambda: L.check(lib.ipdm_conv_first(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), out.data_ptr(), stats.data_ptr(), N, H, W, C, 0, st), "first")
last = lambda: L.check(lib.ipdm_conv_last(a16.data_ptr(), w9.data_ptr(), bias.data_ptr(), sig.data_ptr(), lab.data_ptr(), o1.data_ptr(), ws.data_ptr(), N, H, W, C, st), "last")
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(first); print("conv_first %.4f ms  %.0f GB/s (f32 out)" % (ms, out.numel() * 4 / ms / 1e6))
ms = t(last); print("conv_last  %.4f ms  %.0f GB/s (f16 in)" % (ms, a16.numel() * 2 / ms / 1e6))
# reference check of end_conv against torch on the same f16 inputs
ref = torch.nn.functional.conv2d(a16.float().permute(0, 3, 1, 2), w9.t().reshape(1, C, 3, 3).contiguous(), bias, padding=1)[:, 0] / sig[0]
print("conv_last rel err %.3g" % float((o1 - ref).norm() / ref.norm()))
ref1 = torch.nn.functional.conv2d(x[:, None], w1.reshape(C, 1, 3, 3), b1, padding=1).permute(0, 2, 3, 1)
print("conv_first rel err %.3g" % float((out - ref1).norm() / ref1.norm()))
