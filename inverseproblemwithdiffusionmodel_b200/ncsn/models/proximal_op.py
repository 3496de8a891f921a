"""Mirror of `ncsn/models/proximal_op.py`: data-consistency ("proximal") operators.

`L2Penalty` keeps the reference's behaviour -- `num_steps` plain SGD steps (lr 0.05) from x = z on
    0.5*mean_b sum|x-z|^2 + 0.5*(alpha/lamda)*mean(sum_{dims 1,2,3}|Ax-y|^2)
(proximal_op.py:19-51) -- but evaluates the gradient in closed form instead of building an autograd
graph: grad = (x-z)/B + (alpha/lamda)/D * A^H(Ax - y), with D = Nc*W for the 5-D SENSE output (the sum
runs over (B,C,H), the mean over (Nc,W)) and D = B for a 4-D single-coil operator.
No loss value is printed (the reference's `print(loss.item())` is a host sync, not a result).
"""
import warnings

import torch

from ..linear_transforms import LinearTransform, i2k_complex, k2i_complex, _as_c64
from ..linear_transforms.undersampling_fourier import RandomUndersamplingFourier, SENSE
from ... import _lib

SGD_LR = 5e-2  # proximal_op.py:38


class Proximal(object):
    def __init__(self, lin_tfm: LinearTransform):
        self.lin_tfm = lin_tfm

    def __call__(self, *args, **kwargs):
        pass


def _axpy(a, b, s):
    """a + s*b on complex64 CUDA tensors through the library kernel."""
    out = torch.empty_like(a)
    _lib.check(_lib.lib().ipdm_caxpy(out.data_ptr(), a.data_ptr(), b.data_ptr(), float(s), a.numel(), _lib.stream()), "caxpy")
    return out


def l2_kappa(lin_tfm, z, alpha, lamda):
    """lr * (alpha/lamda) / D  -- the factor in front of A^H(Az - y) after one SGD step."""
    if isinstance(lin_tfm, SENSE):
        D = lin_tfm.sens_maps.shape[0] * z.shape[-1]
    else:
        D = z.shape[0]
    return SGD_LR * (float(alpha) / float(lamda)) / D


class L2Penalty(Proximal):
    def __call__(self, z, y, alpha, lamda, num_steps=1):
        """x <- one (or num_steps) SGD step(s) towards argmin_x 1/2|x - z|^2 + 1/2 alpha/lamda |Ax - y|^2"""
        A = self.lin_tfm
        if not isinstance(A, (SENSE, RandomUndersamplingFourier)):
            raise _lib.IpdmError("L2Penalty: only SENSE and RandomUndersamplingFourier operators are implemented")
        z = _as_c64(z)
        y = _as_c64(y)
        kappa = l2_kappa(A, z, alpha, lamda)
        B = z.shape[0]
        x = z
        for it in range(num_steps):
            resid = _axpy(A(x), y, -1.0)                       # A x - y  (A x is masked; so is every y in use)
            g = A.conj_op(_mask_kspace(A, resid))
            nxt = _axpy(x, g, -kappa)
            if it > 0:
                nxt = _axpy(nxt, _axpy(x, z, -1.0), -SGD_LR / B)  # the (x - z)/B term vanishes on the first step
            x = nxt
        return x.detach()

    @torch.no_grad()
    def check_solution(self, x_sol, z, y, alpha, lamda):
        warnings.warn("For testing only, don't use this in iterations.")
        b = z + alpha / lamda * self.lin_tfm.conj_op(y)
        lhs = x_sol + alpha / lamda * self.lin_tfm.conj_op(self.lin_tfm(x_sol))
        return (torch.abs(lhs - b) ** 2).sum(dim=(1, 2, 3)).mean()


def _mask_kspace(A, S):
    """The true adjoint of `mask * i2k` re-applies the mask; `conj_op` does not (quirk Q3), so the
    gradient path multiplies explicitly (a no-op when y is already masked).  In place on S."""
    m, frames = A.device_mask(S.device)
    H, W = S.shape[-2:]
    # images are ordered (..., b, c): image index % frames == b for C == 1
    _lib.check(_lib.lib().ipdm_kspace_combine(S.data_ptr(), None, m.data_ptr(), frames, 0.0, 2,
                                              S.numel() // (H * W), H, W, _lib.stream()), "mask k-space")
    return S


class Constrained(Proximal):
    """Proximal operator from Yang et al (MRI): k-space projection (proximal_op.py:62-69)."""

    def __call__(self, X: torch.Tensor, S: torch.Tensor, lamda: float):
        return self.lin_tfm.projection(X, S, lamda)


class SingleCoil(Proximal):
    def __init__(self, lin_tfm: RandomUndersamplingFourier):
        super(SingleCoil, self).__init__(lin_tfm)
        assert isinstance(self.lin_tfm, RandomUndersamplingFourier), "only supporting RandomUnversamplingFourier"

    def __call__(self, z, y, alpha, lamda):
        """x = F' diag(1 / (1 + alpha * M_ii)) F (z + alpha F'y)   (proximal_op.py:77-94)"""
        alpha = float(alpha) / float(lamda)
        z = _as_c64(z)
        m, frames = self.lin_tfm.device_mask(z.device)
        x = _axpy(z, k2i_complex(y), alpha)
        K = i2k_complex(x)
        H, W = K.shape[-2:]
        _lib.check(_lib.lib().ipdm_kspace_combine(K.data_ptr(), None, m.data_ptr(), frames, alpha, 0,
                                                  K.numel() // (H * W), H, W, _lib.stream()), "SingleCoil")
        return k2i_complex(K)

    @torch.no_grad()
    def check_solution(self, x_out, z, y, alpha, lamda):
        warnings.warn("For testing only, don't use this in iterations.")
        alpha = alpha / lamda
        lhs = x_out + alpha * self.lin_tfm.conj_op(self.lin_tfm(x_out))
        rhs = alpha * self.lin_tfm.conj_op(y) + z
        return (torch.abs(lhs - rhs) ** 2).sum(dim=(1, 2, 3)).mean()


def get_proximal(proximal_name: str):
    assert proximal_name in ["L2Penalty", "Constrained", "SingleCoil"]
    return {"L2Penalty": L2Penalty, "Constrained": Constrained, "SingleCoil": SingleCoil}[proximal_name]
