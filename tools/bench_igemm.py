"""Per-shape timing of ipdm_conv_igemm (CUDA events, 5 reps after 2 warm-ups): TFLOP/s and HBM GB/s per epilogue mode."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
if len(sys.argv) > 2:      # A/B: ipdm_debug_option(2, v): residual L2 prefetch on (1) / off (0)
    L.ipdm_debug_option(2, int(sys.argv[2]))
shapes = [  # (H, Cin, Cout, taps, dil)
    (256, 128, 128, 9, 1), (128, 256, 256, 9, 1), (64, 256, 256, 9, 1), (32, 512, 512, 9, 1), (128, 128, 128, 9, 1),
    (256, 128, 256, 9, 1), (32, 256, 256, 9, 1), (32, 512, 512, 9, 4), (256, 128, 256, 1, 1)]
modes = {"f16elu": dict(o32=False, o16=True, res=False, st=False, flags=1),
         "f32+stats": dict(o32=True, o16=False, res=False, st=True, flags=0),
         "res+f32+f16elu": dict(o32=True, o16=True, res=True, st=False, flags=1),
         # 16-bit residual stream (what the ngf-128 network launches since round 2)
         "t16: res16+raw16+f16elu": dict(o32=False, o16=True, res=False, st=False, flags=1, res16=True, raw16=True),
         "t16: raw16+stats": dict(o32=False, o16=False, res=False, st=True, flags=0, raw16=True),
         "t16: res16+raw16": dict(o32=False, o16=False, res=False, st=False, flags=0, res16=True, raw16=True)}
if os.environ.get("IGEMM_ALL_MODES"):     # which stream costs what: every combination of the three epilogue streams, first shape only
    shapes = shapes[:1]
    modes.update({"f32": dict(o32=True, o16=False, res=False, st=False, flags=0),
                  "f32+f16elu": dict(o32=True, o16=True, res=False, st=False, flags=1),
                  "res+f32": dict(o32=True, o16=False, res=True, st=False, flags=0),
                  "res+f16elu": dict(o32=False, o16=True, res=True, st=False, flags=1)})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (H, Cin, Cout, taps, dil) in shapes:
    x16 = torch.randn(N, H, H, Cin, device=dev).half()
    w16 = (torch.randn(Cout, taps, Cin, device=dev) / (taps * Cin) ** 0.5).half()
    for name, m in modes.items():
        o32 = torch.empty(N, H, H, Cout, device=dev) if m["o32"] else None
        o16 = torch.empty(N, H, H, Cout, device=dev, dtype=torch.float16) if m["o16"] else None
        res = torch.randn(N, H, H, Cout, device=dev) if m["res"] else None
        st = torch.zeros(N, Cout, 2, device=dev, dtype=torch.float64) if m["st"] else None
        res16 = torch.randn(N, H, H, Cout, device=dev).half() if m.get("res16") else None
        raw16 = torch.empty(N, H, H, Cout, device=dev, dtype=torch.float16) if m.get("raw16") else None
        d = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, _lib.ptr(res), _lib.ptr(o32), _lib.ptr(o16), _lib.ptr(st),
                          N, H, H, Cin, Cout, taps, dil, m["flags"], 0, 0, _lib.ptr(res16), _lib.ptr(raw16))
        for _ in range(2):
            _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            _lib.check(L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flop = 2.0 * N * H * H * Cout * Cin * taps
        byt = N * H * H * (Cin * 2 + Cout * (4 * m["o32"] + 2 * m["o16"] + 4 * m["res"] + 2 * bool(m.get("res16")) + 2 * bool(m.get("raw16"))))
        print(json.dumps({"H": H, "Cin": Cin, "Cout": Cout, "taps": taps, "dil": dil, "mode": name, "ms": round(ms, 4),
                          "tflops": round(flop / ms / 1e9, 1), "hbm_gbs": round(byt / ms / 1e6, 1)}), flush=True)
        del o32, o16, res, st, res16, raw16
