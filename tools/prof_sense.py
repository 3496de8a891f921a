import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import parity_cases as C
nc, n, B, R = 4, 256, 64, 40
A = C.SENSE("exp", nc, R, 1 / 64, (1, n, n), 0)
A.random_under_fourier.mask = C.keep_center_mask(n, R, 1 / 64, seed=0)
x = torch.randn(B, 1, n, n, dtype=torch.complex64, device="cuda")
for _ in range(3):
    S = A(x)
torch.cuda.synchronize()
print("ok")
