"""Mirror of the annealed-Langevin part of `sde/sampling.py`: the 'ald' corrector
(`AnnealedLangevinDynamics.update_fn`, sde/sampling.py:290-324), on the fused update kernel.

Only what the north-star path names is here (the 'ald' corrector, plus the 'langevin' corrector that shares its
update); predictors and the ODE sampler are out of scope.
"""
import torch

from .. import _lib

_CORRECTORS = {}


def register_corrector(cls=None, *, name=None):
    def _register(c):
        key = c.__name__ if name is None else name
        if key in _CORRECTORS:
            raise ValueError(f'Already registered model with name: {key}')
        _CORRECTORS[key] = c
        return c
    return _register if cls is None else _register(cls)


def get_corrector(name):
    return _CORRECTORS[name]


class Corrector:
    """The abstract class for a corrector algorithm (sde/sampling.py:167-190)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps


@register_corrector(name='langevin')
class LangevinCorrector(Corrector):
    """step = (snr * mean_b|noise_b| / mean_b|grad_b|)^2 * 2 * alpha, the same for every sample
    (sde/sampling.py:258-287).  The two batch-mean norms are device-side reductions (no host sync); the update
    itself is the fused Langevin kernel with a per-sample step."""

    def update_fn(self, x, t, noise_fn=None):
        _lib.require_cuda(x, t)
        sde = self.sde
        if hasattr(sde, "alphas"):
            timestep = (t * (sde.N - 1) / sde.T).long()
            alpha = sde.alphas.to(t.device)[timestep]
        else:
            alpha = torch.ones_like(t)
        x = x.detach().float().contiguous().clone()
        x_mean = torch.empty_like(x)
        per = x[0].numel()
        L = _lib.lib()
        for i in range(self.n_steps):
            grad = self.score_fn(x, t).contiguous()
            noise = (torch.randn_like(x) if noise_fn is None else noise_fn(x.shape).to(x.device, torch.float32)).contiguous()
            grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
            noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
            step = ((self.snr * noise_norm / grad_norm) ** 2 * 2 * alpha).float().contiguous()
            _lib.check(L.ipdm_langevin_update(x.data_ptr(), grad.data_ptr(), noise.data_ptr(), x_mean.data_ptr(), x.numel(), None,
                                              None, None, step.data_ptr(), per, None, _lib.stream()), "langevin corrector")
        return x, x_mean


@register_corrector(name='ald')
class AnnealedLangevinDynamics(Corrector):
    """step = (snr*std_t)^2 * 2 * alpha_t ;  x_mean = x + step*score ;  x = x_mean + sqrt(2 step)*noise.
    `sde` needs `.marginal_prob(x, t)[1]` (and `.alphas`, `.N`, `.T` for VP-type SDEs, detected by attribute)."""

    def update_fn(self, x, t, noise_fn=None, seed=None):
        """seed None: every call draws a fresh seed from torch's global generator (the reference draws fresh
        torch.randn_like noise at every call, sde/sampling.py:316); pass a seed to pin the stream of ONE call."""
        _lib.require_cuda(x, t)
        seed = (0 if noise_fn is not None else _lib.fresh_seed()) if seed is None else int(seed)
        sde = self.sde
        if hasattr(sde, "alphas"):
            timestep = (t * (sde.N - 1) / sde.T).long()
            alpha = sde.alphas.to(t.device)[timestep]
        else:
            alpha = torch.ones_like(t)
        std = sde.marginal_prob(x, t)[1]
        x = x.detach().float().contiguous().clone()
        x_mean = torch.empty_like(x)
        per = x[0].numel()
        L = _lib.lib()
        for i in range(self.n_steps):
            grad = self.score_fn(x, t).contiguous()
            step = ((self.snr * std) ** 2 * 2 * alpha).float().contiguous()
            noise = None if noise_fn is None else noise_fn(x.shape).to(x.device, torch.float32).contiguous()
            _lib.check(L.ipdm_langevin_update(x.data_ptr(), grad.data_ptr(), _lib.ptr(noise), x_mean.data_ptr(), x.numel(), None,
                                              None, None, step.data_ptr(), per, _lib.rng(seed, i, None, per), _lib.stream()), "ald corrector")
        return x, x_mean
