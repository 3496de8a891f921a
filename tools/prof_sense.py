"""One launch of each SENSE kernel at a sweep point (for ncu): python tools/prof_sense.py [coils size batch R]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import parity_cases as C
from inverseproblemwithdiffusionmodel_b200 import _lib
L = _lib.lib()
if os.environ.get("IPDM_L2_FETCH"):      # experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes)
    _lib.check(L.ipdm_debug_option(4, int(os.environ["IPDM_L2_FETCH"])))
nc, n, B, R = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (4, 256, 64, 40)
A = C.SENSE("exp", nc, R, 1 / 64, (1, n, n), 0)
A.random_under_fourier.mask = C.keep_center_mask(n, R, 1 / 64, seed=0)
dev = torch.device("cuda")
x = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)
state = torch.randn(2, B, n, n, device=dev); grad = torch.randn_like(state); bvec = torch.randn_like(state)
mre, mim = A.device_maps(dev); m, frames = A.device_mask(dev)
sc = _lib.AldScalars(0.1, 0.4, 0.01, 1.0)
reps = int(os.environ.get("REPS", "2"))
for _ in range(reps):
    S = A(x)
    A.conj_op_masked(S)
    A.conj_op(S)
    _lib.check(L.ipdm_ald_sense_step_plan(A.device_plan(dev, n).handle, state.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), mre.data_ptr(), None,
                                          nc, B, n, sc, None, None, _lib.rng(1, 0), _lib.stream()))
torch.cuda.synchronize()
print("ok")
