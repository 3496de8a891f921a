"""Oracle (test infrastructure): centred FFTs, sampling masks, coil maps, SENSE, proximal steps.

Restates `ncsn/linear_transforms/__init__.py`, `ncsn/linear_transforms/undersampling_fourier.py`
and `ncsn/models/proximal_op.py` of the reference as plain functions on torch CPU tensors.
"""
import numpy as np
import torch


# --------------------------------------------------------------------------- centred orthonormal DFT
def i2k(x: torch.Tensor) -> torch.Tensor:
    """image -> k-space, centred, orthonormal, over the last two axes.

    Reference: `i2k_complex`, ncsn/linear_transforms/__init__.py:36-45
    (cast to complex64, ifftshift, fftn(norm="ortho"), fftshift)."""
    x = x.to(torch.complex64)
    axes = (-2, -1)
    return torch.fft.fftshift(torch.fft.fft2(torch.fft.ifftshift(x, dim=axes), dim=axes, norm="ortho"), dim=axes)


def k2i(s: torch.Tensor) -> torch.Tensor:
    """k-space -> image. Reference: `k2i_complex`, ncsn/linear_transforms/__init__.py:48-57."""
    s = s.to(torch.complex64)
    axes = (-2, -1)
    return torch.fft.fftshift(torch.fft.ifft2(torch.fft.ifftshift(s, dim=axes), dim=axes, norm="ortho"), dim=axes)


# --------------------------------------------------------------------------- masks
def variable_density_masks(T, N, sw=0.3, sm=0.7, sa=0.045, n_candidates=1000, dev=0.01, seed=None):
    """Variable-density column masks, bool, shape (T,1,N) (or (1,N) when T == 1).

    Reference: `generate_mask`, ncsn/linear_transforms/__init__.py:60-76: Bernoulli candidates with
    p = exp(-|x|/sw)*sm+sa, the two centre lines forced on, keep candidates whose sampling rate is
    within `dev` of the mean, draw T of them with replacement.  Seeded through numpy's global RNG
    exactly as the reference does."""
    np.random.seed(seed)
    grid = np.linspace(-1.0, 1.0, N)
    prob = np.exp(-np.abs(grid) / sw) * sm + sa
    cand = np.random.rand(N, n_candidates) <= prob[:, None]
    mid = N // 2
    cand[mid - 1:mid + 1, :] = True
    rate = cand.mean(axis=0)
    keep = np.abs(rate - cand.mean()) < dev
    pool = cand[:, keep]
    pick = np.random.choice(pool.shape[1], T)
    out = pool[:, pick].T
    if T == 1:
        return torch.tensor(out[0:1, :])
    return torch.tensor(out[:, None, :])


def live_sense_mask(W, seed):
    """The mask the reference's `RandomUndersamplingFourier` actually builds today: T=24 frames,
    the "R = 16" parameter set, shape (24,1,1,W); ctor args R / center_lines_frac are ignored.
    Reference: undersampling_fourier.py:63-75 (quirk Q1)."""
    torch.random.manual_seed(seed)
    return variable_density_masks(24, W, sw=0.07926, sm=0.42, sa=0.02, seed=seed).unsqueeze(1)


def keep_center_mask(W, R, center_lines_frac, seed):
    """The retired keep-centre rule kept in comments at undersampling_fourier.py:50-61: per-column
    Bernoulli(1/R) from torch's RNG, a centred window of int(W*frac) lines forced on. Float (1,1,W)."""
    torch.random.manual_seed(seed)
    mask = (torch.rand(1, 1, W) <= 1.0 / R).float()
    win = int(W * center_lines_frac)
    start = W // 2 - win // 2
    mask[..., start:start + win] = 1.0
    return mask


# --------------------------------------------------------------------------- coil maps
def exp_coil_maps(num_coils, H, W, seed):
    """Real float64 coil sensitivities (Nc,H,W), exp(-dist/(2l)) around a random anchor, normalised
    so that sum_c |s_c|^2 == 1.  Reference: SENSE.__init__/_generate_sens_map,
    undersampling_fourier.py:101-138.  The coordinate grid is mgrid[0:W,0:H] flattened and reshaped
    to (H,W) just like the reference (only self-consistent for square images, quirk Q13)."""
    maps = []
    for i in range(num_coils):
        np.random.seed(None if seed is None else seed + i)
        ah, aw = np.random.choice(H), np.random.choice(W)
        ww, hh = np.mgrid[0:W, 0:H]
        pts = np.stack([ww.ravel(), hh.ravel()], axis=1).astype(np.float64)
        dist = np.sqrt(((pts - np.array([[ah, aw]], dtype=np.float64)) ** 2).sum(axis=1))
        ell = dist.max() / 2
        maps.append(torch.exp(-torch.tensor(dist.reshape(H, W)) / (2 * ell)))
    maps = torch.stack(maps, dim=0)
    return maps / torch.sqrt((maps.abs() ** 2).sum(dim=0))


# --------------------------------------------------------------------------- operators
def undersampled_fourier(x, mask):
    """A1 x = mask * i2k(x). Reference: RandomUndersamplingFourier.__call__, undersampling_fourier.py:77-82."""
    return mask.to(x.device) * i2k(x)


def sense_forward(x, maps, mask):
    """S[c] = mask * i2k(s_c * x): (B,C,H,W) -> (Nc,B,C,H,W).
    Reference: SENSE.__call__, undersampling_fourier.py:140-150 (float64 maps promote the product
    to complex128 before i2k rounds it to complex64, quirk Q4)."""
    return torch.stack([undersampled_fourier(maps[c] * x, mask) for c in range(maps.shape[0])], dim=0)


def sense_adjoint(S, maps):
    """sum_c conj(s_c) * k2i(S[c]) -- no mask (quirk Q3). Reference: SENSE.conj_op, :152-160."""
    out = torch.zeros(S.shape[1:], dtype=S.dtype, device=S.device)
    for c in range(S.shape[0]):
        out += maps[c].conj() * k2i(S[c])
    return out


def sense_ssos(S):
    """sqrt(sum_c |k2i(S[c])|^2), float32. Reference: SENSE.SSOS, :162-170."""
    acc = torch.zeros(S.shape[1:], dtype=torch.float32, device=S.device)
    for c in range(S.shape[0]):
        acc += k2i(S[c]).abs() ** 2
    return acc.sqrt()


def log_lh_grad(fwd, adj, x, s, lamda=1.0):
    """-lamda * A^H(Ax - s). Reference: LinearTransform.log_lh_grad, linear_transforms/__init__.py:26-33."""
    return -adj(fwd(x) - s) * lamda


def fourier_projection(x, s, mask, lamda):
    """k-space blend. Reference: RandomUndersamplingFourier.projection, undersampling_fourier.py:89-97."""
    mask = mask.to(x.device)
    kx = i2k(x)
    return k2i(lamda * s + (1 - lamda) * mask * kx + (1 - mask) * kx)


# --------------------------------------------------------------------------- proximal steps
def l2_prox_sgd(fwd, z, y, alpha, lamda, num_steps=1, lr=5e-2):
    """The reference's `L2Penalty`: `num_steps` plain SGD steps (lr 0.05) from x = z on
    0.5*mean_b sum|x-z|^2 + 0.5*(alpha/lamda)*mean(sum_{dims 1,2,3}|A x - y|^2), by autograd.
    Reference: proximal_op.py:19-51."""
    x = z.clone().requires_grad_(True)
    with torch.enable_grad():
        for _ in range(num_steps):
            fit = 0.5 * ((x - z).abs() ** 2).sum(dim=(1, 2, 3)).mean()
            data = 0.5 * alpha / lamda * ((fwd(x) - y).abs() ** 2).sum(dim=(1, 2, 3)).mean()
            (g,) = torch.autograd.grad(fit + data, x)
            x = (x - lr * g).detach().requires_grad_(True)
    return x.detach()


def l2_prox_sense_closed_form(z, y, maps, mask, alpha, lamda, lr=5e-2):
    """One-step closed form of the above for the 5-D SENSE output:
    the sum runs over (B,C,H) and the mean over (Nc,W), hence
    x = z - lr*(alpha/lamda)/(Nc*W) * A^H(A z - y)."""
    Nc, W = maps.shape[0], z.shape[-1]
    kappa = lr * (alpha / lamda) / (Nc * W)
    resid = sense_forward(z, maps, mask) - y
    return z - kappa * sense_adjoint(mask.to(z.device) * resid, maps)


def single_coil_prox(z, y, mask, alpha, lamda):
    """Exact single-coil prox x = k2i(i2k(z + a*k2i(y)) / (1 + a*mask)), a = alpha/lamda.
    Reference: SingleCoil.__call__, proximal_op.py:72-94."""
    a = alpha / lamda
    mask = mask.to(z.device)
    return k2i(i2k(z + a * k2i(y)) * (1 / (1 + mask * a)))


def prox_residual(fwd, adj, x, z, y, alpha, lamda):
    """|| x + a A^H A x - (z + a A^H y) ||^2 summed per sample, batch mean.
    Reference: check_solution, proximal_op.py:53-59 / 96-104."""
    a = alpha / lamda
    lhs = x + a * adj(fwd(x))
    rhs = z + a * adj(y)
    return ((lhs - rhs).abs() ** 2).sum(dim=(1, 2, 3)).mean()
