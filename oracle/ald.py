"""Oracle (test infrastructure): the annealed-Langevin samplers of the reference as plain loops.

Restates `ncsn/models/__init__.py` (get_sigmas, anneal_Langevin_dynamics),
`ncsn/models/ALD_optimizers.py` (ALDOptimizer, ALDInvSegProximalRealImag, ALD2DTime) and the
'ald' corrector of `sde/sampling.py`.  Noise is drawn through a caller-supplied `draw(shape)`
callable so tests can inject the very tensors the reference consumed (order documented per
function); the default draws from torch's global CPU generator exactly like `torch.randn_like`.
"""
import numpy as np
import torch


def _default_draw(shape):
    return torch.randn(*shape)


def geometric_sigmas(sigma_begin, sigma_end, num_classes):
    """exp(linspace(log s1, log sL, L)) in float64 numpy, then .float().
    Reference: get_sigmas, ncsn/models/__init__.py:10-38 (geometric branch)."""
    return torch.tensor(np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), num_classes))).float()


def uniform_sigmas(sigma_begin, sigma_end, num_classes):
    """Reference: get_sigmas, uniform branch, ncsn/models/__init__.py:18-21,32-35."""
    return torch.tensor(np.linspace(sigma_begin, sigma_end, num_classes)).float()


def langevin_update(x, grad, noise, step_size):
    """x + step*grad + noise*sqrt(2*step). Reference: ALD_optimizers.py:117,239,241,444-445."""
    return x + step_size * grad + noise * torch.sqrt(step_size * 2)


def ald_unconditional(score, x, sigmas, n_steps_each, step_lr, denoise=True, draw=_default_draw):
    """Reference: ALDOptimizer.__call__ (ALD_optimizers.py:66-137) == anneal_Langevin_dynamics
    (ncsn/models/__init__.py:40-82). One noise draw per inner step. Returns the final x."""
    B = x.shape[0]
    for c in range(len(sigmas)):
        labels = torch.full((B,), c, dtype=torch.long)
        step = step_lr * (sigmas[c] / sigmas[-1]) ** 2
        for _ in range(n_steps_each):
            g = score(x, labels)
            x = langevin_update(x, g, draw(x.shape), step)
    if denoise:
        last = torch.full((B,), len(sigmas) - 1, dtype=torch.long)
        x = x + sigmas[-1] ** 2 * score(x, last)
    return x


def ald_sense_real_imag(score, measurement, sigmas, n_steps_each, step_lr, lr_scaled, adjoint, prox,
                        denoise=True, draw=_default_draw, trace=None):
    """cfg 2/3 sampler with guidance weight 0.  Reference: ALDInvSegProximalRealImag.__call__ and
    post_processing, ALD_optimizers.py:172-327.  x0 = A^H y; per step: score on real and imag
    separately, noise drawn in the order real, imag (:238-241), Langevin update on each part, then
    `prox(z, y, step_lr*lr_scaled, 1.)` on the recombined complex image (alpha is the *unscaled*
    step_lr, quirk Q6); final denoise on each part.  `trace`, if a list, receives x after every step."""
    x = adjoint(measurement)
    xr, xi = x.real, x.imag
    B = x.shape[0]
    for c in range(len(sigmas)):
        labels = torch.full((B,), c, dtype=torch.long)
        step = step_lr * (sigmas[c] / sigmas[-1]) ** 2
        for _ in range(n_steps_each):
            gr = score(xr, labels)
            gi = score(xi, labels)
            xr = langevin_update(xr, gr, draw(xr.shape), step)
            xi = langevin_update(xi, gi, draw(xi.shape), step)
            z = prox(xr + 1j * xi, measurement, step_lr * lr_scaled, 1.0)
            xr, xi = z.real, z.imag
            if trace is not None:
                trace.append((xr + 1j * xi).clone())
    if denoise:
        last = torch.full((B,), len(sigmas) - 1, dtype=torch.long)
        xr = xr + sigmas[-1] ** 2 * score(xr, last)
        xi = xi + sigmas[-1] ** 2 * score(xi, last)
    return xr + 1j * xi


def temporal_tv_grad(x, lamda):
    """-lamda * D^T sign(D x) with circular forward differences along dim 1.
    Reference: FiniteDiff (dims=1), ncsn/linear_transforms/finite_diff.py:7-35."""
    s = torch.sign(torch.roll(x, -1, 1) - x)
    return -lamda * (torch.roll(s, 1, 1) - s)


def ald_2dtime(score, measurement, sigmas, n_steps_each, step_lr, lr_scaled, adjoint, prox,
               mode_T="none", lamda_T=1.0, draw=_default_draw):
    """cfg 4 sampler for mode_T in {"none","tv"}.  Reference: ALD2DTime.__call__ / init_x_mod /
    spatial_step / temporal_step / proximal_step, ALD_optimizers.py:351-554.  measurement is
    (Nc,B,T,C,H,W); both noises are drawn before either update (:442-443); no final denoise."""
    Nc, B, T, C, H, W = measurement.shape
    y = measurement.reshape(Nc, B * T, C, H, W)
    x = adjoint(y)
    for c in range(len(sigmas)):
        labels = torch.full((B * T,), c, dtype=torch.long)
        step = step_lr * (sigmas[c] / sigmas[-1]) ** 2
        for _ in range(n_steps_each):
            xr, xi = x.real, x.imag
            gr = score(xr, labels)
            gi = score(xi, labels)
            nr = draw(xr.shape)
            ni = draw(xi.shape)
            xr = langevin_update(xr, gr, nr, step)
            xi = langevin_update(xi, gi, ni, step)
            if "tv" in mode_T:
                xr5 = xr.reshape(B, T, C, H, W)
                xi5 = xi.reshape(B, T, C, H, W)
                xr = (xr5 + temporal_tv_grad(xr5, lamda_T)).reshape(B * T, C, H, W)
                xi = (xi5 + temporal_tv_grad(xi5, lamda_T)).reshape(B * T, C, H, W)
            x = prox(xr + 1j * xi, y, step_lr * lr_scaled, 1.0)
    return x.reshape(B, T, C, H, W)


def sde_ald_corrector(score_fn, x, t, std, snr, n_steps, alpha=None, draw=_default_draw):
    """'ald' corrector of the vendored score_sde sampler: step = (snr*std)^2 * 2 * alpha;
    x_mean = x + step*score; x = x_mean + noise*sqrt(2*step).
    Reference: AnnealedLangevinDynamics.update_fn, sde/sampling.py:303-324."""
    alpha = torch.ones_like(t) if alpha is None else alpha
    x_mean = x
    for _ in range(n_steps):
        g = score_fn(x, t)
        noise = draw(x.shape)
        step = (snr * std) ** 2 * 2 * alpha
        x_mean = x + step[:, None, None, None] * g
        x = x_mean + noise * torch.sqrt(step * 2)[:, None, None, None]
    return x, x_mean


# --------------------------------------------------------------------------- posterior statistics / metrics
def posterior_stats(recons):
    """mean / population-std of magnitude and phase over the chain axis.
    Reference: helpers/visualizations.py:93-95,117-142 (numpy mean/std, ddof=0)."""
    mag, ph = recons.abs(), torch.angle(recons)
    return {"mag_mean": mag.mean(0), "mag_std": mag.std(0, unbiased=False),
            "phase_mean": ph.mean(0), "phase_std": ph.std(0, unbiased=False)}


def nrmse(recon_mag, orig_mag):
    """skimage normalized_root_mse(recon, orig, 'euclidean') with the reference's argument order,
    i.e. normalised by the *reconstruction's* norm (quirk Q9). Reference: helpers/metrics.py:70-74."""
    return float(torch.sqrt(((recon_mag - orig_mag) ** 2).mean()) / torch.sqrt((recon_mag ** 2).mean()))


def ssim(a, b, data_range=None, win=7, K1=0.01, K2=0.03):
    """skimage-default structural similarity for one 2-D image pair (7x7 uniform window, sample
    covariance, mean over the valid interior).  The reference calls skimage (helpers/metrics.py:55-68),
    which is absent here; `data_range` must be given explicitly and identically on both sides."""
    import torch.nn.functional as F

    a = a.double()[None, None]
    b = b.double()[None, None]
    if data_range is None:
        data_range = float(b.max() - b.min())
    k = torch.ones(1, 1, win, win, dtype=torch.float64) / (win * win)
    n = win * win
    cov_norm = n / (n - 1)
    ua, ub = F.conv2d(a, k), F.conv2d(b, k)
    uaa, ubb, uab = F.conv2d(a * a, k), F.conv2d(b * b, k), F.conv2d(a * b, k)
    va, vb, vab = cov_norm * (uaa - ua * ua), cov_norm * (ubb - ub * ub), cov_norm * (uab - ua * ub)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ua * ub + C1) * (2 * vab + C2)) / ((ua ** 2 + ub ** 2 + C1) * (va + vb + C2))
    return float(S.mean())
