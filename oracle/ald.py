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
                        denoise=True, draw=_default_draw, trace=None, guide=None):
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
            if guide is not None:   # guide(x_part, c) -> seg-likelihood gradient already scaled by weight_c / sigma_c (:272-286)
                gr = gr + guide(xr, c)
                gi = gi + guide(xi, c)
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


def remap_sigmas_T(sigmas, sigmas_T):
    """Temporal schedule nearest-interpolated onto the tail of the spatial one; -1 = skip the temporal step.
    Reference: ALD2DTime.__init__, ALD_optimizers.py:342-345 (quirk Q14: this vector also REPLACES the temporal
    network's own sigmas)."""
    import torch.nn.functional as F
    n = int((sigmas <= sigmas_T[0]).sum())
    out = torch.ones_like(sigmas) * (-1)
    if n > 0:
        out[-n:] = F.interpolate(sigmas_T.view(1, 1, -1), n, mode="nearest").squeeze()
    return out


def fold_patches(x, k):
    """(N, T, H, W) -> (N*H/k*W/k, k*k, T). Reference: reshape_temporal_dim 'forward', helpers/utils.py:330-345."""
    N, T, H, W = x.shape
    return x.reshape(N, T, H // k, k, W // k, k).permute(0, 2, 4, 3, 5, 1).reshape(-1, k * k, T)


def unfold_patches(p, k, H, W):
    """Inverse of fold_patches. Reference: reshape_temporal_dim 'backward', helpers/utils.py:347-359."""
    T = p.shape[-1]
    return p.reshape(-1, H // k, W // k, k, k, T).permute(0, 5, 1, 3, 2, 4).reshape(-1, T, H, W)


def ald_2dtime(score, measurement, sigmas, n_steps_each, step_lr, lr_scaled, adjoint, prox,
               mode_T="none", lamda_T=1.0, draw=_default_draw, score_T=None, sigmas_T=None, win=8, random_shift=False):
    """cfg 4 sampler for mode_T in {"none","tv","diffusion1d"}.  Reference: ALD2DTime.__call__ / init_x_mod /
    spatial_step / temporal_step / proximal_step, ALD_optimizers.py:351-554.  measurement is
    (Nc,B,T,C,H,W); both noises are drawn before either update (:442-443); no final denoise.
    "diffusion1d": score_T(patches (B', win*win, T), labels) with `sigmas_T` already remapped (remap_sigmas_T),
    skipped where it is -1, optional np.random roll of the frames before folding (:463-502)."""
    import numpy as np
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
            elif "diffusion1d" in mode_T and float(sigmas_T[c]) != -1:
                vr = xr.reshape(B, T, C, H, W).permute(0, 2, 1, 3, 4).reshape(-1, T, H, W)
                vi = xi.reshape(B, T, C, H, W).permute(0, 2, 1, 3, 4).reshape(-1, T, H, W)
                if random_shift:
                    sh = tuple(np.random.randint(0, win, (2,)).tolist())
                    vr, vi = torch.roll(vr, sh, (-2, -1)), torch.roll(vi, sh, (-2, -1))
                pr, pi = fold_patches(vr, win), fold_patches(vi, win)
                lab = torch.full((pr.shape[0],), c, dtype=torch.long)
                step_T = step_lr * (sigmas_T[c] / sigmas_T[-1]) ** 2 * lamda_T
                sr, si = score_T(pr, lab), score_T(pi, lab)
                n1, n2 = draw(pr.shape), draw(pi.shape)
                pr = langevin_update(pr, sr, n1, step_T)
                pi = langevin_update(pi, si, n2, step_T)
                vr, vi = unfold_patches(pr, win, H, W), unfold_patches(pi, win, H, W)
                if random_shift:
                    back = tuple(-v for v in sh)
                    vr, vi = torch.roll(vr, back, (-2, -1)), torch.roll(vi, back, (-2, -1))
                xr = vr.reshape(B, C, T, H, W).permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
                xi = vi.reshape(B, C, T, H, W).permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
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


def sde_langevin_corrector(score_fn, x, t, snr, n_steps, alpha=None, draw=_default_draw):
    """'langevin' corrector: step = (snr*mean|noise|/mean|grad|)^2 * 2 * alpha (batch-mean per-sample norms).
    Reference: LangevinCorrector.update_fn, sde/sampling.py:268-287."""
    alpha = torch.ones_like(t) if alpha is None else alpha
    x_mean = x
    for _ in range(n_steps):
        g = score_fn(x, t)
        noise = draw(x.shape)
        gn = torch.norm(g.reshape(g.shape[0], -1), dim=-1).mean()
        nn_ = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
        step = (snr * nn_ / gn) ** 2 * 2 * alpha
        x_mean = x + step[:, None, None, None] * g
        x = x_mean + torch.sqrt(step * 2)[:, None, None, None] * noise
    return x, x_mean


def seg_guidance_grad(seg, x, label, mode="full"):
    """d/dx sum log softmax(seg(x))[label].  Reference: compute_seg_grad, ncsn/models/__init__.py:197-215."""
    with torch.enable_grad():
        X = x.detach().clone().requires_grad_(True)
        prob = torch.softmax(seg(X), dim=1)
        torch.log(torch.gather(prob, dim=1, index=label)).sum().backward()
        g = X.grad
    return g * label if mode == "FG" else g


def map_sense(score, x0, y, fwd, adj, lamda, lr, n_iters, betas=(0.5, 0.5)):
    """MAP baseline: Adam on x with grad = -A^H(Ax - y) + lamda*(score(Re x) + i score(Im x)), label 1 throughout.
    Reference: MAPOptimizer._step, ncsn/models/MAP_optimizers.py:97-115 (torch.optim.Adam on the complex tensor)."""
    x = x0.clone()
    opt = torch.optim.Adam([x], lr=lr, betas=betas)
    labels = torch.ones(x.shape[0]).long()
    for _ in range(n_iters):
        grad = -adj(fwd(x) - y) + lamda * torch.complex(score(x.real, labels), score(x.imag, labels))
        opt.zero_grad()
        x.grad = -grad
        opt.step()
    return x


def map_2dtime_tv(score, x0, y6, fwd, adj, lr, n_iters, prior_weight, w_S, w_T, betas=(0.5, 0.5), score_T=None, win=8):
    """2D+time MAP: separate Adam optimisers on the real and imaginary parts; temporal term = TV (score_T None) or
    the learned patch prior with label 1 (mode_T = "diffusion1d", no roll).  Reference: MAPOptimizer2DTime,
    ncsn/models/MAP_optimizers.py:154-306."""
    B, T, C, H, W = x0.shape
    xr, xi = x0.real.clone(), x0.imag.clone()
    o_r = torch.optim.Adam([xr], lr=lr, betas=betas)
    o_i = torch.optim.Adam([xi], lr=lr, betas=betas)
    y = y6.reshape(y6.shape[0], B * T, C, H, W)

    def grad_of(x):
        xf = x.reshape(B * T, C, H, W)
        g_data = (-adj(fwd(xf) - y)).reshape(B, T, C, H, W)
        labels = torch.ones(B * T).long()
        g_S = torch.complex(score(xf.real, labels), score(xf.imag, labels)).reshape(B, T, C, H, W)
        if score_T is None:
            g_T = torch.complex(temporal_tv_grad(x.real, 1.0), temporal_tv_grad(x.imag, 1.0))
        else:
            v = x.permute(0, 2, 1, 3, 4).reshape(-1, T, H, W)
            pr, pi = fold_patches(v.real, win), fold_patches(v.imag, win)
            lab = torch.ones(pr.shape[0]).long()
            g = torch.complex(unfold_patches(score_T(pr, lab), win, H, W), unfold_patches(score_T(pi, lab), win, H, W))
            g_T = g.reshape(B, C, T, H, W).permute(0, 2, 1, 3, 4)
        return g_data + prior_weight * (w_S * g_S + w_T * g_T)

    # Quirk of the reference: x_real / x_imag are VIEWS of the initial x, so in iteration 0 the imaginary closure
    # sees the already-updated real part; from iteration 1 on self.x is a fresh tensor that the in-place optimiser
    # steps no longer touch, so both closures evaluate the gradient at the x of the iteration start (:175-176,206-233)
    for it in range(n_iters):
        x_start = torch.complex(xr, xi)
        o_r.zero_grad()
        xr.grad = -grad_of(x_start).real.contiguous()
        o_r.step()
        o_i.zero_grad()
        xi.grad = -grad_of(torch.complex(xr, xi) if it == 0 else x_start).imag.contiguous()
        o_i.step()
    return torch.complex(xr, xi)


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
