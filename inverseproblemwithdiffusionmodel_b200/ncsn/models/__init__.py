"""Mirror of `ncsn/models/__init__.py` (hot-path part): `get_sigmas` and the functional
`anneal_Langevin_dynamics`."""
import numpy as np
import torch

from ... import _lib


def get_sigmas(config, mode="unconditioned"):
    """Noise schedule (reference ncsn/models/__init__.py:10-38): geometric or uniform, computed in
    float64 numpy then cast to float32, on `config.device`."""
    assert mode in ("unconditioned", "recons")
    src = config.recons if mode == "recons" else config.model
    if src.sigma_dist == 'geometric':
        vals = np.exp(np.linspace(np.log(src.sigma_begin), np.log(src.sigma_end), src.num_classes))
    elif src.sigma_dist == 'uniform':
        vals = np.linspace(src.sigma_begin, src.sigma_end, src.num_classes)
    else:
        raise NotImplementedError('sigma distribution not supported')
    return torch.tensor(vals).float().to(config.device)


def ald_schedule(sigmas, n_steps_each, step_lr, kappa=0.0):
    """Per-inner-step scalars of the whole chain as a float32 (L*n_steps_each, 4) tensor
    (step, sqrt(2*step), kappa, sigma) -- the device-side schedule the captured step graph indexes.
    step = step_lr * (sigma / sigma_L)^2 in float32 like the reference (ALD_optimizers.py:101)."""
    sig = sigmas.detach().float().cpu()
    step = (torch.tensor(step_lr, dtype=torch.float32) * (sig / sig[-1]) ** 2)
    rows = torch.stack([step, torch.sqrt(step * 2), torch.full_like(step, float(kappa)), sig], dim=1)
    return rows.repeat_interleave(n_steps_each, dim=0).contiguous()


def langevin_update_(x, grad, step, noise=None, seed=0, rng_step=0, x_mean=None, chain_ids=None):
    """In place x += step*grad + sqrt(2*step)*noise on float32 CUDA tensors (B, ...).  noise None: in-kernel Philox,
    sample i drawing the stream of chain chain_ids[i] (default i)."""
    _lib.require_cuda(x, grad)
    sc = _lib.AldScalars(float(step), float(torch.sqrt(torch.tensor(float(step), dtype=torch.float32) * 2)), 0.0, 0.0)
    per = x[0].numel() if x.dim() > 1 else 0
    _lib.check(_lib.lib().ipdm_langevin_update(x.data_ptr(), grad.data_ptr(), _lib.ptr(noise), _lib.ptr(x_mean), x.numel(),
                                               sc, None, None, None, 0, _lib.rng(seed, rng_step, chain_ids, per), _lib.stream()),
               "langevin_update")
    return x


@torch.no_grad()
def anneal_Langevin_dynamics(x_mod, scorenet, sigmas, n_steps_each=200, step_lr=0.000008,
                             final_only=False, verbose=False, denoise=True, noise_fn=None, seed=None):
    """Reference ncsn/models/__init__.py:40-82.  `noise_fn(shape) -> Tensor` injects noise (parity
    tests); otherwise noise comes from the in-kernel Philox stream `seed` (None: a fresh seed from torch's
    global generator per call, like the reference's randn_like)."""
    _lib.require_cuda(x_mod)
    seed = (0 if noise_fn is not None else _lib.fresh_seed()) if seed is None else int(seed)
    x_mod = x_mod.detach().float().contiguous().clone()
    images = []
    k = 0
    for c, sigma in enumerate(sigmas):
        labels = torch.full((x_mod.shape[0],), c, dtype=torch.long, device=x_mod.device)
        step_size = step_lr * (sigma / sigmas[-1]) ** 2
        for s in range(n_steps_each):
            grad = scorenet(x_mod, labels)
            noise = None if noise_fn is None else noise_fn(x_mod.shape).to(x_mod.device, torch.float32).contiguous()
            langevin_update_(x_mod, grad.contiguous(), step_size, noise=noise, seed=seed, rng_step=k)
            k += 1
            if not final_only:
                images.append(x_mod.to('cpu'))
    if denoise:
        last = torch.full((x_mod.shape[0],), len(sigmas) - 1, dtype=torch.long, device=x_mod.device)
        x_mod = x_mod + sigmas[-1] ** 2 * scorenet(x_mod, last)
        images.append(x_mod.to('cpu'))
    if final_only:
        return [x_mod.to('cpu')]
    return images
