"""Mirror of `ncsn/linear_transforms/undersampling_fourier.py`: `RandomUndersamplingFourier`
and the multi-coil `SENSE` operator, on the fused kernels of csrc/sense.cu.

Construction (mask and coil-map generation) is host-side numpy, bit-identical to the reference;
`__call__` / `conj_op` / `SSOS` / `projection` run on CUDA tensors only.  The quirks that define
parity are kept on purpose: the live mask ignores `R` / `center_lines_frac` and is
(24,1,1,W) (Q1/Q2), `conj_op` applies no mask (Q3), the coil maps are real float64 (Q4).  The mask
is a plain attribute that callers may overwrite (e.g. with a (1,1,W) keep-centre mask).
"""
import contextlib
import warnings

import numpy as np
import torch

from . import LinearTransform, generate_mask, workspace, _as_c64, fft2c
from ... import _lib


def keep_center_mask(W, R, center_lines_frac, seed):
    """The retired rule kept in comments in the reference (:50-61): Bernoulli(1/R) columns from torch's
    RNG plus a centred window of int(W*frac) lines; float32 (1,1,W).  Used for the "R = 40" benchmark."""
    torch.random.manual_seed(seed)
    mask = (torch.rand(1, 1, W) <= 1 / R).float()
    win = int(W * center_lines_frac)
    lo = W // 2 - win // 2
    mask[..., lo:lo + win] = 1.
    return mask


class _MaskCache:
    """uint8 [frames][W] device copy of a broadcastable column mask and its compiled plans (one per image height),
    rebuilt when the attribute changes."""

    def __init__(self):
        self.key = None
        self.dev = None
        self.host = None
        self.frames = 1
        self.plans = {}

    def get(self, mask, device):
        key = (id(mask), mask._version, device)
        if key != self.key:
            W = mask.shape[-1]
            if mask.numel() % W != 0 or (mask.dim() >= 2 and mask.shape[-2] != 1 and mask.numel() != W):
                raise _lib.IpdmError(f"mask of shape {tuple(mask.shape)} is not a column mask (…,1,W)")
            flat = (mask.reshape(-1, W) != 0).to(torch.uint8)
            self.host = flat.cpu().contiguous()
            self.dev = flat.to(device).contiguous()
            self.frames = flat.shape[0]
            self.key = key
            self.plans = {}
        return self.dev, self.frames

    def plan(self, mask, device, H):
        """`_lib.SensePlan` of this mask for images of H rows on `device` (created on first use, outside graph capture)."""
        self.get(mask, device)
        pl = self.plans.get(H)
        if pl is None:
            with (torch.cuda.device(device) if torch.device(device).type == "cuda" else contextlib.nullcontext()):
                pl = _lib.SensePlan(self.host.numpy(), H, self.host.shape[-1])
            self.plans[H] = pl
        return pl


def _frames_vs_batch(frames, X):
    """Reference broadcasting of a (T,1,1,W) mask against (B,C,H,W): B must be 1 or T (Q2)."""
    B = X.shape[0]
    if frames == 1 or frames == B:
        return X
    if B == 1:
        return X.expand(frames, *X.shape[1:])
    raise RuntimeError(f"The size of tensor a ({frames}) must match the size of tensor b ({B}) at non-singleton dimension 0")


class RandomUndersamplingFourier(LinearTransform):
    def __init__(self, R, center_lines_frac, in_shape, seed=None):
        """in_shape: (C, H, W)"""
        self.R = R
        self.center_lines_frac = center_lines_frac
        self.in_shape = in_shape
        self.seed = seed
        self.mask = self._generate_mask()
        self._mc = _MaskCache()

    def _generate_mask(self):
        # (T, 1, 1, W) with T = 24 and the "R = 16" parameters -- exactly what the reference builds (:63-75)
        torch.random.manual_seed(self.seed)
        W = self.in_shape[-1]
        return generate_mask(24, W, sw=0.07926, sm=0.42, sa=0.02, seed=self.seed).unsqueeze(1)

    def device_mask(self, device):
        return self._mc.get(self.mask, device)

    def device_plan(self, device, H):
        return self._mc.plan(self.mask, device, H)

    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        X = _as_c64(X)
        m, frames = self.device_mask(X.device)
        X = _frames_vs_batch(frames, X).contiguous()
        if frames > 1 and X.shape[1] != 1:
            raise _lib.IpdmError("per-frame masks need C == 1")
        return fft2c(X, inverse=False, plan=self.device_plan(X.device, X.shape[-2]))

    def conj_op(self, S: torch.Tensor) -> torch.Tensor:
        return fft2c(S, inverse=True)

    def projection(self, X: torch.Tensor, S: torch.Tensor, lamda: float) -> torch.Tensor:
        # k2i(lamda*S + (1-lamda)*mask*i2k(X) + (1-mask)*i2k(X))   (reference :89-97)
        X = _as_c64(X)
        S = _as_c64(S)
        m, frames = self.device_mask(X.device)
        K = fft2c(X, inverse=False)
        H, W = K.shape[-2:]
        _lib.check(_lib.lib().ipdm_kspace_combine(K.data_ptr(), S.data_ptr(), m.data_ptr(), frames, float(lamda), 1,
                                                  K.numel() // (H * W), H, W, _lib.stream()), "projection")
        return fft2c(K, inverse=True)


class SENSE(LinearTransform):
    def __init__(self, sens_type, num_sens, R, center_lines_frac, in_shape, seed):
        assert sens_type in ["exp"]
        self.random_under_fourier = RandomUndersamplingFourier(R, center_lines_frac, in_shape, seed)
        maps = []
        for i in range(num_sens):
            s = self.random_under_fourier.seed
            maps.append(self._generate_sens_map(sens_type, None if s is None else s + i))
        maps = torch.stack(maps, dim=0)                      # (num_sens, H, W) float64
        self.sens_maps = maps / torch.sqrt((maps.abs() ** 2).sum(dim=0))
        energy = (self.sens_maps.abs() ** 2).sum(dim=0)
        assert torch.allclose(energy, torch.ones_like(energy))
        self._maps_key = None
        self._maps_dev = None

    def _generate_sens_map(self, sens_type, seed=0, **kwargs):
        # exp(-||p - p0|| / (2 l)), p0 drawn with np.random.seed(seed), l = max distance / 2 (reference :119-138)
        H, W = self.random_under_fourier.in_shape[-2:]
        np.random.seed(seed)
        p0 = np.array([np.random.choice(H), np.random.choice(W)], dtype=np.float64)
        ww, hh = np.mgrid[0:W, 0:H]
        pts = np.stack([ww.flatten(), hh.flatten()], axis=1).astype(np.float64)
        dist = np.sqrt(((pts - p0[None, :]) ** 2).sum(axis=1))
        ell = kwargs.get("l", dist.max() / 2)
        return torch.exp(-torch.tensor(dist.reshape((H, W))) / (2 * ell))

    # ---- device constants ---------------------------------------------------------------------
    def device_maps(self, device):
        """(maps_re, maps_im or None) as contiguous float32 device tensors."""
        m = self.sens_maps
        key = (id(m), m._version, device)
        if key != self._maps_key:
            if torch.is_complex(m):
                self._maps_dev = (m.real.float().contiguous().to(device), m.imag.float().contiguous().to(device))
            else:
                self._maps_dev = (m.float().contiguous().to(device), None)
            self._maps_key = key
        return self._maps_dev

    def device_mask(self, device):
        return self.random_under_fourier.device_mask(device)

    def device_plan(self, device, H):
        return self.random_under_fourier.device_plan(device, H)

    # ---- operator -------------------------------------------------------------------------------
    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        """X: (B, C, H, W) -> (num_sens, B, C, H, W)   (reference :140-150)"""
        X = _as_c64(X)
        m, frames = self.device_mask(X.device)
        X = _frames_vs_batch(frames, X).contiguous()
        if frames > 1 and X.shape[1] != 1:
            raise _lib.IpdmError("per-frame masks need C == 1")
        mre, mim = self.device_maps(X.device)
        Nc = mre.shape[0]
        H, W = X.shape[-2:]
        batch = X.numel() // (H * W)
        out = torch.empty((Nc,) + tuple(X.shape), dtype=torch.complex64, device=X.device)
        L = _lib.lib()
        ws = workspace(X.device, L.ipdm_sense_workspace_bytes(Nc, batch, H, W))
        plan = self.device_plan(X.device, H)
        _lib.check(L.ipdm_sense_forward_plan(plan.handle, X.data_ptr(), mre.data_ptr(), _lib.ptr(mim), out.data_ptr(),
                                             Nc, batch, ws.data_ptr(), _lib.stream()), "SENSE.__call__")
        return out

    def _adjoint(self, S, ssos, masked):
        S = _as_c64(S)
        mre, mim = self.device_maps(S.device)
        Nc = S.shape[0]
        if Nc != mre.shape[0]:
            raise _lib.IpdmError(f"SENSE: got {Nc} coil images for {mre.shape[0]} coil maps")
        H, W = S.shape[-2:]
        batch = S[0].numel() // (H * W)
        out = torch.empty(S.shape[1:], dtype=torch.float32 if ssos else torch.complex64, device=S.device)
        L = _lib.lib()
        ws = workspace(S.device, L.ipdm_sense_workspace_bytes(Nc, batch, H, W))
        if masked:
            plan = self.device_plan(S.device, H)
            _lib.check(L.ipdm_sense_adjoint_plan(plan.handle, S.data_ptr(), mre.data_ptr(), _lib.ptr(mim), out.data_ptr(),
                                                 Nc, batch, 1 if ssos else 0, ws.data_ptr(), _lib.stream()), "SENSE.conj_op_masked")
        else:
            _lib.check(L.ipdm_sense_adjoint(S.data_ptr(), mre.data_ptr(), _lib.ptr(mim), None, 1, out.data_ptr(),
                                            Nc, batch, H, W, 1 if ssos else 0, ws.data_ptr(), _lib.stream()), "SENSE.conj_op")
        return out

    def conj_op(self, S: torch.Tensor) -> torch.Tensor:
        """S: (num_sens, B, C, H, W) -> sum_c conj(s_c) k2i(S_c), no mask (reference :152-160)"""
        return self._adjoint(S, ssos=False, masked=False)

    def conj_op_masked(self, S: torch.Tensor) -> torch.Tensor:
        """conj_op for inputs that are already zero off the mask (y, A x, A x - y): skips those columns."""
        return self._adjoint(S, ssos=False, masked=True)

    def SSOS(self, S: torch.Tensor) -> torch.Tensor:
        """sqrt(sum_c |k2i(S_c)|^2), float32 (reference :162-170)"""
        return self._adjoint(S, ssos=True, masked=False)

    def projection(self, X: torch.Tensor, S: torch.Tensor, lamda: float) -> torch.Tensor:
        warnings.warn("Not implemented!")
        return X
