"""Mirror of `ncsn/linear_transforms/__init__.py` of the reference: the `LinearTransform` interface,
the centred orthonormal FFT helpers and the variable-density mask generator.

The FFT helpers run the fused row/column kernels of csrc/sense.cu (no cuFFT, no fftshift copies).
"""
import abc

import numpy as np
import torch

from ... import _lib


class LinearTransform(abc.ABC):
    """All inputs: (B, C, H, W).  Same four-method interface as the reference
    (ncsn/linear_transforms/__init__.py:6-33)."""

    @abc.abstractmethod
    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        return X

    @abc.abstractmethod
    def conj_op(self, S: torch.Tensor) -> torch.Tensor:
        return S

    @abc.abstractmethod
    def projection(self, X: torch.Tensor, S: torch.Tensor, lamda: float) -> torch.Tensor:
        return X

    def log_lh_grad(self, X: torch.Tensor, S: torch.Tensor, lamda: float = 1.) -> torch.Tensor:
        """grad = -lamda * A'(Ax - s)   (reference :26-33)"""
        diff = self(X) - S
        return -self.conj_op(diff) * lamda


_WS = {}


def workspace(device, nbytes):
    """Per-device scratch for the transposed FFT intermediate; grows, never shrinks."""
    key = (device.type, device.index)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def _as_c64(X):
    _lib.require_cuda(X)
    if X.dtype != torch.complex64:
        X = X.to(torch.complex64)
    return X.contiguous()


def fft2c(X, inverse=False, mask_u8=None, mask_frames=1, plan=None):
    """Centred orthonormal 2-D DFT of the last two axes of a CUDA tensor (any leading shape).  A column mask is
    given as a compiled plan (`_lib.SensePlan`) or as a raw u8 [frames][W] device tensor."""
    X = _as_c64(X)
    H, W = X.shape[-2:]
    batch = X.numel() // (H * W)
    out = torch.empty_like(X)
    L = _lib.lib()
    ws = workspace(X.device, L.ipdm_sense_workspace_bytes(1, batch, H, W))
    if plan is not None:
        if inverse:
            _lib.check(L.ipdm_sense_adjoint_plan(plan.handle, X.data_ptr(), None, None, out.data_ptr(), 1, batch, 0, ws.data_ptr(),
                                                 _lib.stream()), "k2i_complex")
        else:
            _lib.check(L.ipdm_sense_forward_plan(plan.handle, X.data_ptr(), None, None, out.data_ptr(), 1, batch, ws.data_ptr(),
                                                 _lib.stream()), "i2k_complex")
    elif inverse:
        _lib.check(L.ipdm_sense_adjoint(X.data_ptr(), None, None, _lib.ptr(mask_u8), mask_frames, out.data_ptr(),
                                        1, batch, H, W, 0, ws.data_ptr(), _lib.stream()), "k2i_complex")
    else:
        _lib.check(L.ipdm_sense_forward(X.data_ptr(), None, None, _lib.ptr(mask_u8), mask_frames, out.data_ptr(),
                                        1, batch, H, W, ws.data_ptr(), _lib.stream()), "i2k_complex")
    return out


def i2k_complex(X):
    """X: (B, C, D, H, W) or (B, C, H, W) -> centred k-space, complex64 (reference :36-45)."""
    return fft2c(X, inverse=False)


def k2i_complex(X):
    """centred k-space -> image (reference :48-57)."""
    return fft2c(X, inverse=True)


def generate_mask(T: int, N: int, sw=0.3, sm=0.7, sa=0.045, T_max=1000, dev=0.01, seed=None):
    """Variable-density column masks; host-side setup, same RNG stream as the reference (:60-76):
    bool (T, 1, N), or (1, N) when T == 1."""
    np.random.seed(seed)
    pos = np.linspace(-1, 1, N)
    keep_prob = sm * np.exp(-np.abs(pos) / sw) + sa
    trial = np.random.rand(N, T_max) <= keep_prob[:, None]
    trial[N // 2 - 1:N // 2 + 1, :] = True
    ok = np.abs(trial.mean(axis=0) - trial.mean()) < dev
    trial = trial[:, ok]
    chosen = np.random.choice(trial.shape[1], T)
    masks = np.ascontiguousarray(trial[:, chosen].T)
    if T == 1:
        return torch.tensor(masks[0:1, :])
    return torch.tensor(masks[:, None, :])
