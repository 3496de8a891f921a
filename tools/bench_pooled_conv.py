"""ConvMeanPool two ways (CUDA events, 10 reps after 3): the 3x3 convolution at full resolution with the 2x2 mean in the
epilogue, and the equivalent 4x4 stride-2 convolution = 3x3 over the space-to-depth operand with 16 of 36 blocks (tap mask)."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
dev = "cuda"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timeit(d, reps=10):
    for _ in range(3):
        _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

shapes = [(28, 256, 128, 256), (28, 128, 256, 256), (28, 64, 256, 256)]
if len(sys.argv) > 1 and sys.argv[1] == "small":     # one chain at 256^2 (cfg2-B1), 48 frames at 128^2 (cfg 4), 16 images at 28^2 (cfg 1)
    shapes = [(2, 256, 128, 256), (2, 128, 256, 256), (2, 64, 256, 256), (48, 128, 128, 256), (48, 64, 256, 256), (48, 32, 256, 256), (16, 28, 128, 256)]
for (N, H, Cin, Cout) in shapes:
    Ho = H // 2
    x = torch.randn(N, H, H, Cin, device=dev).half()
    xs = torch.randn(N, Ho, Ho, 4 * Cin, device=dev).half()
    w = (torch.randn(Cout, 9, Cin, device=dev) / 30).half()
    ws = (torch.randn(Cout, 9, 4 * Cin, device=dev) / 60).half()
    res = torch.randn(N, Ho, Ho, Cout, device=dev).half()
    raw = torch.empty_like(res); elu = torch.empty_like(res)
    st = torch.zeros(N, Cout, 2, device=dev, dtype=torch.float64)
    d1 = _lib.ConvDesc(x.data_ptr(), w.data_ptr(), None, None, None, elu.data_ptr(), st.data_ptr(), N, H, H, Cin, Cout, 9, 1,
                       _lib.CONV_F16_ELU | _lib.CONV_POOL2, 0, 0, res.data_ptr(), raw.data_ptr())
    d2 = _lib.ConvDesc(xs.data_ptr(), ws.data_ptr(), None, None, None, elu.data_ptr(), st.data_ptr(), N, Ho, Ho, 4 * Cin, Cout, 9, 1,
                       _lib.CONV_F16_ELU, 0, 0, res.data_ptr(), raw.data_ptr())
    if 4 * Cin <= 1024:
        for kc in range(4 * Cin // 64):
            par = (kc * 64) // Cin
            rows = (1, 2) if (par >> 1) == 0 else (0, 1)
            cols = (1, 2) if (par & 1) == 0 else (0, 1)
            d2.tap_mask[kc] = sum(1 << (3 * r + c) for r in rows for c in cols)
    t1, t2 = timeit(d1), timeit(d2)
    d3 = _lib.ConvDesc(xs.data_ptr(), ws.data_ptr(), None, None, None, elu.data_ptr(), st.data_ptr(), N, Ho, Ho, 4 * Cin, Cout, 9, 1,
                       _lib.CONV_F16_ELU, 0, 0, res.data_ptr(), raw.data_ptr())
    t3 = timeit(d3)
    print(json.dumps({"N": N, "H": H, "Cin": Cin, "Cout": Cout, "pooled_epilogue_ms": round(t1, 4), "s2d_masked_ms": round(t2, 4),
                      "s2d_unmasked_ms": round(t3, 4)}), flush=True)
