"""Mirror of the MAP baselines of `ncsn/models/MAP_optimizers.py` on the B200 operators (a "next" row of
SURVEY.md 8f): gradient ascent on the log-posterior with the score network as the prior gradient,
    grad = -A^H(A x - y) + lamda * (score(Re x) + i score(Im x)),   optimiser step on -grad.
The score network and `log_lh_grad` run on the library kernels; the optimiser itself is `torch.optim` exactly
as in the reference (Adam with betas (0.5, 0.5) by default, MAP_optimizers.py:78-81).  Logging (TensorBoard in
the reference) is optional: pass `logger=None` to skip it.
"""
import torch

from ..linear_transforms import LinearTransform
from ..linear_transforms.finite_diff import FiniteDiff
from ... import _lib


class _NullLogger:
    def add_scalar(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass


class MAPOptimizer(object):
    def __init__(self, x_init: torch.Tensor, measurement: torch.Tensor, scorenet, linear_tfm: LinearTransform, lamda,
                 config, logger=None, device=None, opt_class=None, opt_params=None):
        """x_init: (B, C, H, W) complex CUDA tensor (updated in place); reference :58-81"""
        self.x_init = x_init
        self.measurement = measurement
        self.scorenet = scorenet
        self.linear_tfm = linear_tfm
        self.lamda = lamda
        self.config = config
        self.device = x_init.device if device is None else device
        self.logger = logger if logger is not None else _NullLogger()
        self.verbose = logger is not None
        self.lr = self.config.MAP.lr
        if opt_class is None:
            opt_class = torch.optim.Adam
            opt_params = {"betas": (0.5, 0.5)}
        self.opt = opt_class([self.x_init], lr=self.lr, **(opt_params or {}))

    @torch.no_grad()
    def __call__(self):
        _lib.require_cuda(self.x_init)
        x = self.x_init
        for it in range(self.config.MAP.n_iters):
            x = self._step(x, it)
            if self.verbose:   # a host sync per iteration, only when somebody listens
                err = 0.5 * (torch.norm(self.linear_tfm(x) - self.measurement) ** 2)
                self.logger.add_scalar("data_error", err.item(), global_step=it)
        return x

    def _step(self, x, it):
        grad_data = self.linear_tfm.log_lh_grad(x, self.measurement, 1.)
        labels = torch.ones(x.shape[0], device=x.device).long()          # the reference always uses level 1 (:101)
        grad_prior = torch.complex(self.scorenet(torch.real(x), labels), self.scorenet(torch.imag(x), labels))
        grad = grad_data + self.lamda * grad_prior
        self.opt.zero_grad()
        self.x_init.grad = -grad
        self.opt.step()
        return x


class SENSEMAP(MAPOptimizer):
    pass


class MAPOptimizer2DTime(object):
    def __init__(self, x_init, measurement, scorenet_S, scorenet_T, linear_tfm, logger, params):
        """x_init: (B, T, C, H, W) complex; measurement (num_sens, B, T, C, H, W); params as in the reference
        (:155-176): lr, opt_class, num_iters, prior_weight, spatial_step_weight, temporal_step_weight, mode_T, ..."""
        self.params = params
        self.x = x_init
        self.x_real = torch.real(self.x).contiguous()
        self.x_imag = torch.imag(self.x).contiguous()
        self.measurement = measurement
        self.scorenet_S, self.scorenet_T = scorenet_S, scorenet_T
        self.linear_tfm = linear_tfm
        self.logger = logger if logger is not None else _NullLogger()
        oc, op = params["opt_class"], params.get("opt_params", {})
        self.opt_real = oc([self.x_real], lr=params["lr"], **op)
        self.opt_imag = oc([self.x_imag], lr=params["lr"], **op)
        self.finite_diff = None

    def _grad(self):
        p = self.params
        grad_data = self.data_step()
        grad_S = self.spatial_step()
        grad_T = self.temporal_step(mode_T=p["mode_T"], if_random_shift=p.get("if_random_shift", False))
        return grad_data + p["prior_weight"] * (p["spatial_step_weight"] * grad_S + p["temporal_step_weight"] * grad_T)

    @torch.no_grad()
    def __call__(self):
        for it in range(self.params["num_iters"]):
            # Reference quirk (MAP_optimizers.py:175-176,206-233): x_real / x_imag are views of the INITIAL x, so in
            # iteration 0 the imaginary closure sees the updated real part, while from iteration 1 on `self.x` is a
            # fresh tensor the in-place optimiser steps do not touch -- both closures then see the x of the iteration
            # start (one gradient evaluation serves both).
            g = self._grad()
            self.opt_real.zero_grad()
            self.x_real.grad = -torch.real(g).contiguous()
            self.opt_real.step()
            if it == 0:
                self.x = torch.complex(self.x_real, self.x_imag)
                g = self._grad()
            self.opt_imag.zero_grad()
            self.x_imag.grad = -torch.imag(g).contiguous()
            self.opt_imag.step()
            self.x = torch.complex(self.x_real, self.x_imag)
        return self.get_reconstruction()

    def data_step(self):
        B, T, C, H, W = self.x.shape
        x = self.x.reshape(B * T, C, H, W)
        y = self.measurement.reshape(self.measurement.shape[0], B * T, C, H, W)
        return self.linear_tfm.log_lh_grad(x, y).reshape(B, T, C, H, W)

    def spatial_step(self):
        B, T, C, H, W = self.x.shape
        x = self.x.reshape(B * T, C, H, W)
        labels = torch.ones(x.shape[0], device=x.device).long()
        g = torch.complex(self.scorenet_S(torch.real(x), labels), self.scorenet_S(torch.imag(x), labels))
        return g.reshape(B, T, C, H, W)

    def temporal_step(self, mode_T="tv", if_random_shift=False):
        if mode_T == "diffusion1d":
            return self._temporal_diffusion(if_random_shift)
        if mode_T != "tv":
            raise _lib.IpdmError(f"mode_T={mode_T!r}: expected 'tv' or 'diffusion1d'")
        if self.finite_diff is None:
            self.finite_diff = FiniteDiff(dims=1)
        return torch.complex(self.finite_diff.log_lh_grad(torch.real(self.x).contiguous()),
                             self.finite_diff.log_lh_grad(torch.imag(self.x).contiguous()))

    def _temporal_diffusion(self, if_random_shift):
        """score of the learned temporal prior on k x k x T patches, label 1 (reference :284-306): fold (with the optional
        np.random roll) -> scorenet_T on the real and imaginary patches -> unfold the gradient."""
        import numpy as np
        B, T, C, H, W = self.x.shape
        if C != 1:
            raise _lib.IpdmError("MAPOptimizer2DTime: C must be 1")
        k = int(self.params["win_size"])
        L = _lib.lib()
        planar = torch.stack([torch.real(self.x), torch.imag(self.x)]).reshape(2, B * T, H, W).to(torch.float32).contiguous()
        _lib.require_cuda(planar)
        P2 = 2 * B * (H // k) * (W // k)
        vol = torch.empty(P2, k, T, k, dtype=torch.float32, device=planar.device)
        gvol = torch.empty_like(vol)
        sh, sw = (tuple(np.random.randint(0, k, (2,)).tolist()) if if_random_shift else (0, 0))
        _lib.check(L.ipdm_patch_fold(planar.data_ptr(), vol.data_ptr(), B, T, H, W, k, sh, sw, 0, _lib.stream()), "patch_fold")
        labels = torch.ones(P2, dtype=torch.long, device=planar.device)
        if hasattr(self.scorenet_T, "forward_into"):
            self.scorenet_T.forward_into(vol, labels, gvol)
        else:
            gvol.copy_(self.scorenet_T(vol.permute(0, 1, 3, 2).reshape(P2, k * k, T), labels).reshape(P2, k, k, T).permute(0, 1, 3, 2))
        _lib.check(L.ipdm_patch_fold(planar.data_ptr(), gvol.data_ptr(), B, T, H, W, k, sh, sw, 1, _lib.stream()), "patch_unfold")
        return torch.complex(planar[0], planar[1]).reshape(B, T, C, H, W)

    def get_reconstruction(self):
        return self.x.detach().cpu()
