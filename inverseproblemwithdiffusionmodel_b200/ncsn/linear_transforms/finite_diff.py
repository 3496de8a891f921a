"""Mirror of `ncsn/linear_transforms/finite_diff.py`: circular temporal finite differences.

`log_lh_grad` on a CUDA tensor of shape (B, T, ...) with dims=1 is what `ALD2DTime` uses for its
"tv" temporal step (ALD_optimizers.py:455-462); it is served by the `ipdm_temporal_tv_step` kernel.
"""
from typing import Tuple, Union

import torch

from . import LinearTransform
from ... import _lib


class FiniteDiff(LinearTransform):
    def __init__(self, dims: Union[int, Tuple[int]]):
        self.dims = dims

    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        return torch.roll(X, -1, self.dims) - X

    def conj_op(self, S: torch.Tensor) -> torch.Tensor:
        return torch.roll(S, 1, self.dims) - S

    def projection(self, X, S, lamda):
        return X

    def log_lh_grad(self, X: torch.Tensor, S: torch.Tensor = None, lamda: float = 1) -> torch.Tensor:
        """grad = -lamda * nabla' sign(nabla X)   (reference :29-35); real float32 (B, T, ...) CUDA input."""
        _lib.require_cuda(X)
        if self.dims != 1 or X.dtype != torch.float32:
            raise _lib.IpdmError("FiniteDiff.log_lh_grad: only dims=1 on float32 (B, T, ...) tensors is implemented")
        B, T = X.shape[:2]
        hw = X[0, 0].numel()
        # the kernel works in place on a planar [2][B][T][hw] state; run it on a two-plane copy
        work = torch.stack([X.contiguous(), X.contiguous()], dim=0)
        _lib.check(_lib.lib().ipdm_temporal_tv_step(work.data_ptr(), B, T, hw, float(lamda), _lib.stream()), "temporal_tv_step")
        return work[0] - X
