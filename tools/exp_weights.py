"""TIMING EXPERIMENT: how fast is k_conv_halo when the weight tiles are NOT re-streamed from L2 (debug option 1 = 2:
results are wrong, only the time matters)?  Separates 'L2 -> SM weight traffic' from 'MMA / smem port' as the limiter."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
dev = "cuda"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (N, H, Cin, Cout) in ((28, 256, 128, 128), (28, 128, 256, 256)):
    x16 = torch.randn(N, H, H, Cin, device=dev).half()
    w16 = (torch.randn(Cout, 9, Cin, device=dev) / (9 * Cin) ** 0.5).half()
    o16 = torch.empty(N, H, H, Cout, device=dev, dtype=torch.float16)
    o32 = torch.empty(N, H, H, Cout, device=dev); res = torch.randn(N, H, H, Cout, device=dev)
    for mode, d in (("f16", _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, None, None, o16.data_ptr(), None, N, H, H, Cin, Cout, 9, 1, 1)),
                    ("res", _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, res.data_ptr(), o32.data_ptr(), o16.data_ptr(), None, N, H, H, Cin, Cout, 9, 1, 1))):
        for variant in (0, 2):
            _lib.check(L.ipdm_debug_option(1, variant))
            for _ in range(2):
                _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
            torch.cuda.synchronize(); e0.record()
            for _ in range(8):
                _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 8
            print(json.dumps({"shape": f"{Cin}->{Cout}@{H}", "mode": mode, "weights": "streamed" if variant == 0 else "NOT re-streamed (experiment)",
                              "ms": round(ms, 4), "tflops": round(2.0 * N * H * H * Cout * Cin * 9 / ms / 1e9, 1)}), flush=True)
    _lib.check(L.ipdm_debug_option(1, 0))
