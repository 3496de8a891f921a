"""How fast can HBM go on the residual-mode conv's traffic mix?  ipdm_add_act: 2 fp32 reads + 1 fp32 write of 940 MB each
(2.82 GB, the same bytes as one k_conv_halo<res+f32+f16> launch at 28 x 256^2 x 128), vs 1 read + 1 write kernels."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
dev = "cuda"
n = 28 * 256 * 256 * 128
a = torch.randn(n, device=dev); b = torch.randn(n, device=dev); o = torch.empty(n, device=dev); h = torch.empty(n, device=dev, dtype=torch.float16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: _lib.check(L.ipdm_add_act(a.data_ptr(), b.data_ptr(), o.data_ptr(), n, 0, _lib.stream())))
print(json.dumps({"kernel": "add_act 2r+1w fp32", "ms": round(ms, 4), "gbs": round(12 * n / ms / 1e6, 1)}))
ms = t(lambda: _lib.check(L.ipdm_add_act(a.data_ptr(), b.data_ptr(), a.data_ptr(), n, 0, _lib.stream())))
print(json.dumps({"kernel": "add_act in place (2r+1w, write over a read stream)", "ms": round(ms, 4), "gbs": round(12 * n / ms / 1e6, 1)}))
ms = t(lambda: _lib.check(L.ipdm_act_to_f16(a.data_ptr(), h.data_ptr(), n, 1, _lib.stream())))
print(json.dumps({"kernel": "act_to_f16 1r fp32 + 1w f16", "ms": round(ms, 4), "gbs": round(6 * n / ms / 1e6, 1)}))
ms = t(lambda: o.copy_(a))
print(json.dumps({"kernel": "torch copy 1r+1w fp32", "ms": round(ms, 4), "gbs": round(8 * n / ms / 1e6, 1)}))
