"""Launch one conv configuration a few times (for ncu): python tools/prof_one.py MODE [N H Cin Cout]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
mode = sys.argv[1] if len(sys.argv) > 1 else "res"
N, H, Cin, Cout = (int(a) for a in sys.argv[2:6]) if len(sys.argv) > 5 else (28, 256, 128, 128)
dev = "cuda"
x16 = torch.randn(N, H, H, Cin, device=dev).half()
w16 = (torch.randn(Cout, 9, Cin, device=dev) / (9 * Cin) ** 0.5).half()
o32 = torch.empty(N, H, H, Cout, device=dev) if mode in ("res", "f32") else None
o16 = torch.empty(N, H, H, Cout, device=dev, dtype=torch.float16) if mode in ("res", "f16") else None
res = torch.randn(N, H, H, Cout, device=dev) if mode == "res" else None
if mode == "t16":      # the launch the network issues on the 16-bit residual stream: f16 residual in, f16 result + f16 ELU copy out
    o16 = torch.empty(N, H, H, Cout, device=dev, dtype=torch.float16)
    raw16 = torch.empty_like(o16)
    res16 = torch.randn(N, H, H, Cout, device=dev).half()
    d = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, None, None, o16.data_ptr(), None, N, H, H, Cin, Cout, 9, 1, 1,
                      0, 0, res16.data_ptr(), raw16.data_ptr())
else:
    d = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, _lib.ptr(res), _lib.ptr(o32), _lib.ptr(o16), None, N, H, H, Cin, Cout, 9, 1, 1)
for _ in range(4):
    _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
torch.cuda.synchronize()
print("ok")
