"""cfg-5 microbenchmark: SENSE forward / adjoint / fused ALD step, achieved algorithmic GB/s vs the HBM roofline.
Algorithmic bytes (SURVEY 8d, N = B*H*W): fwd/adj 8N(1+Nc) + 4*Nc*H*W ; fused step 32N + 4*Nc*H*W (Philox noise)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import parity_cases as C
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
if os.environ.get("IPDM_L2_FETCH"):      # experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes)
    _lib.check(L.ipdm_debug_option(4, int(os.environ["IPDM_L2_FETCH"])))
dev = torch.device("cuda")
peak = 6551.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)      # 256 MB that is only ever read


def flush_l2():
    """Cold AND clean L2: write 256 MB (evicts everything), then read another 256 MB (evicts the dirty lines of the write
    sweep, whose write-back would otherwise be charged to the timed kernel: ~126 MB = 20 us of DRAM writes).  FLUSH=write
    keeps the write-only sweep of the earlier rounds."""
    flush.zero_()
    if os.environ.get("FLUSH") != "write":
        flush_rd.sum()

def timeit(fn, reps=5):
    """best of `reps` CUDA-event times of one replay of fn captured in a CUDA graph (no host launch gaps), L2 flushed before each"""
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        flush_l2()                         # evict L2 between timed iterations
        torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

cases = [(4, 256, 1, 40), (4, 256, 14, 40), (4, 256, 64, 40), (8, 256, 16, 16), (32, 128, 64, 4), (4, 512, 16, 40),
         (16, 512, 16, 16), (32, 512, 64, 40), (32, 512, 16, 4), (4, 128, 64, 4)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cases = cases[:4]
for (nc, n, B, R) in cases:
    A = C.SENSE("exp", nc, R, 1 / 64, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, R, 1 / 64, seed=0)
    lines = int(A.random_under_fourier.mask.sum())
    x = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)
    S = A(x)
    N = B * n * n
    bytes_fa = 8 * N * (1 + nc) + 4 * nc * n * n
    t_f = timeit(lambda: A(x))
    t_a = timeit(lambda: A.conj_op(S))
    t_am = timeit(lambda: A.conj_op_masked(S))
    state = torch.randn(2, B, n, n, device=dev); grad = torch.randn_like(state); bvec = torch.randn_like(state)
    mre, mim = A.device_maps(dev); m, frames = A.device_mask(dev)
    sc = _lib.AldScalars(0.1, 0.4, 0.01, 1.0)
    plan = A.device_plan(dev, n)
    step = lambda: _lib.check(L.ipdm_ald_sense_step_plan(plan.handle, state.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), mre.data_ptr(), None,
                                                         nc, B, n, sc, None, None, _lib.rng(1, 0), _lib.stream()))
    t_s = timeit(step)
    bytes_s = 32 * N + 4 * nc * n * n
    noise = torch.randn_like(state)      # the same step with the noise injected (no Philox / Box-Muller in the kernel; 8N more bytes)
    step_inj = lambda: _lib.check(L.ipdm_ald_sense_step_plan(plan.handle, state.data_ptr(), grad.data_ptr(), noise.data_ptr(), bvec.data_ptr(), mre.data_ptr(),
                                                             None, nc, B, n, sc, None, None, _lib.rng(1, 0), _lib.stream()))
    t_si = timeit(step_inj)
    row = {"coils": nc, "size": n, "batch": B, "R": R, "lines": lines, "pruned": plan.pruned, "kspace_MB": round(8 * nc * N / 1e6, 1),
           "step_injected_noise_ms": round(t_si, 4), "step_injected_noise_frac": round((bytes_s + 8 * N) / t_si / 1e6 / peak, 3),
           "fwd_ms": round(t_f, 4), "fwd_GBs": round(bytes_fa / t_f / 1e6, 1), "fwd_frac": round(bytes_fa / t_f / 1e6 / peak, 3),
           "adj_ms": round(t_a, 4), "adj_GBs": round(bytes_fa / t_a / 1e6, 1), "adj_frac": round(bytes_fa / t_a / 1e6 / peak, 3),
           "adj_masked_ms": round(t_am, 4), "adj_masked_GBs": round(bytes_fa / t_am / 1e6, 1),
           "step_ms": round(t_s, 4), "step_GBs": round(bytes_s / t_s / 1e6, 1), "step_frac": round(bytes_s / t_s / 1e6 / peak, 3)}
    print(json.dumps(row), flush=True)
    del x, S, state, grad, bvec
