"""Reconstruction metrics on the device -- drop-in for the reference's `helpers/metrics.py` (numpy + skimage on
the host, helpers/metrics.py:21-137), computed by `ipdm_image_sums` / `ipdm_ssim` / `ipdm_chain_stats_accumulate`
without moving the reconstructions off the GPU.

Same function names, argument order and return conventions as the reference:
`compute_metrics(metric_names, img, img_orig, reduce=None) -> {name: np.ndarray (B,) or scalar}` with
`REGISTERED_METRICS = {"L2", "L1", "SSIM", "NRMSE"}`; inputs are `(B, C, H, W)` / `(C, H, W)` real arrays
(numpy or torch; host inputs are copied to the current CUDA device).  Reference quirks kept on purpose:
NRMSE is normalised by the norm of its FIRST argument, the reconstruction (argument order of
`normalized_root_mse(img, img_orig)`, metrics.py:70-74).  One deliberate difference: skimage
infers SSIM's `data_range` from the dtype (and newer versions refuse to for floats) while the reference leaves
skimage unpinned, so `data_range` is an explicit keyword here (default: max - min of `img_orig`, the value
skimage's documentation recommends).
"""
import numpy as np
import torch

from .. import _lib


def _dev_f32(a):
    t = torch.as_tensor(a)
    if torch.is_complex(t):
        raise _lib.IpdmError("metrics take real arrays: pass np.abs(recons) like the reference does")
    if not t.is_cuda and torch.cuda.is_available():
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    _lib.require_cuda(t)          # no CPU fallback: raises without a GPU
    return t.to(torch.float32).contiguous()


def add_first_channel(img):
    """(C, H, W) -> (1, C, H, W)   (reference :9-18)"""
    if img.ndim == 3:
        return img[None, ...]
    if img.ndim == 4:
        return img
    raise ValueError("Input must have 3 or 4 dimensions.")


def _pair(img, img_orig):
    a, r = add_first_channel(_dev_f32(img)), add_first_channel(_dev_f32(img_orig))
    if r.shape[0] not in (1, a.shape[0]) or r.shape[1:] != a.shape[1:]:
        raise _lib.IpdmError(f"metrics: shapes {tuple(a.shape)} and {tuple(r.shape)} do not match")
    return a, r


def image_sums(img, img_orig):
    """(B, 4) float64 device tensor: sum (a-b)^2, sum a^2, sum b^2, sum |a-b| per image (all of C, H, W)."""
    a, r = _pair(img, img_orig)
    out = torch.empty(a.shape[0], 4, dtype=torch.float64, device=a.device)
    _lib.check(_lib.lib().ipdm_image_sums(a.data_ptr(), r.data_ptr(), out.data_ptr(), a.shape[0], r.shape[0],
                                          a[0].numel(), _lib.stream()), "image_sums")
    return out, a[0].numel()


def mean_squared_error(img, img_orig):
    s, n = image_sums(img, img_orig)
    return (s[:, 0] / n).cpu().numpy()


def MAE(img, img_orig):
    s, n = image_sums(img, img_orig)
    return (s[:, 3] / n).cpu().numpy()


def NRMSE_wrapper(img, img_orig):
    """sqrt(mean (img - orig)^2) / sqrt(mean img^2): 'euclidean' normalisation by the FIRST argument (reference :70-74)."""
    s, _ = image_sums(img, img_orig)
    return torch.sqrt(s[:, 0] / s[:, 1]).cpu().numpy()


def SSIM_wrapper(img, img_orig, data_range=None):
    """skimage structural_similarity defaults per image; multi-channel inputs average the per-channel values
    (channel_axis=0 in the reference, :55-68)."""
    a, r = _pair(img, img_orig)
    B, C, H, W = a.shape
    if data_range is None:
        data_range = float(r.max() - r.min())
    out = torch.empty(B * C, dtype=torch.float64, device=a.device)
    rr = r if r.shape[0] == B else r.expand(B, C, H, W).contiguous()
    _lib.check(_lib.lib().ipdm_ssim(a.data_ptr(), rr.data_ptr(), out.data_ptr(), B * C, B * C, H, W, float(data_range),
                                    _lib.stream()), "ssim")
    return (out.reshape(B, C).mean(1) / ((H - 6) * (W - 6))).cpu().numpy()


REGISTERED_METRICS = {"L2": mean_squared_error, "L1": MAE, "SSIM": SSIM_wrapper, "NRMSE": NRMSE_wrapper}
REGISTERED_REDUCTION = {"mean": np.mean, "sum": np.sum, "max": np.max}


def compute_metrics(metric_names, img, img_orig, reduce=None, **kwargs):
    """img: (B, C, H, W); img_orig: (1 or B, C, H, W).  Returns {name: (B,) array} (or reduced scalars) -- reference :21-45."""
    out = {}
    for name in metric_names:
        assert name in REGISTERED_METRICS
        fn = REGISTERED_METRICS[name]
        vals = fn(img, img_orig, **kwargs) if name == "SSIM" else fn(img, img_orig)
        out[name] = REGISTERED_REDUCTION[reduce](vals) if reduce is not None else vals
    return out


def compute_mean_and_std(imgs):
    """(B, C, H, W) -> real: (mean, std of |.|); complex: (mag_mean, phase_mean, mag_std, phase_std), each (C, H, W),
    population std as numpy's (reference :77-92).  Complex inputs use the chain-statistics kernel."""
    t = torch.as_tensor(imgs)
    assert t.shape[0] > 1
    if not t.is_cuda and torch.cuda.is_available():
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    _lib.require_cuda(t)
    if not torch.is_complex(t):
        t = t.float()
        return t.mean(0), t.abs().std(0, unbiased=False)
    from ..chains import PosteriorStats
    st = PosteriorStats(t[0].numel(), t.device)
    st.add(t.to(torch.complex64).contiguous())
    r = st.finalize(tuple(t.shape[1:]))
    return r["mag_mean"], r["phase_mean"], r["mag_std"], r["phase_std"]


def compute_snr(imgs):
    """20 log10(max / std) of the magnitudes per image (reference :95-102)."""
    t = torch.as_tensor(imgs).abs().float()
    flat = t.reshape(t.shape[0], -1)
    return (20 * torch.log10(flat.max(1).values / flat.std(1, unbiased=False))).cpu().numpy()
