"""Does the relative placement of the residual / f32 out / f16 out tensors matter for the residual-mode conv?
Times k_conv_halo<res+f32+f16> (128->128 @256^2, 28 images) with the three tensors carved out of one arena at byte skews."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
dev = "cuda"
N, H, C = 28, 256, 128
x16 = torch.randn(N, H, H, C, device=dev).half()
w16 = (torch.randn(C, 9, C, device=dev) / (9 * C) ** 0.5).half()
n32, n16 = N * H * H * C * 4, N * H * H * C * 2
arena = torch.empty(2 * n32 + n16 + (64 << 20), dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
base = arena.data_ptr()
base += (-base) % (2 << 20)
for skew in (0, 256, 1024, 4096, 16384, 65536 + 256, 1 << 20, (1 << 20) + 4096 + 256, 3 * (1 << 20) + 12288):
    res_p = base
    o32_p = base + n32 + (-(n32)) % (2 << 20) + skew
    o16_p = o32_p + n32 + (-(n32)) % (2 << 20) + 2 * skew
    d = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, res_p, o32_p, o16_p, None, N, H, H, C, C, 9, 1, 1)
    for _ in range(2):
        _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
    torch.cuda.synchronize(); e0.record()
    for _ in range(8):
        _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    print(json.dumps({"skew_bytes": skew, "ms": round(ms, 4), "hbm_gbs": round((n32 * 2 + n16 * 2) / ms / 1e6, 1)}), flush=True)
