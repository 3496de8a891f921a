"""Mirror of `ncsn/models/ALD_optimizers.py`: the annealed-Langevin samplers.

Same constructors, hooks and return convention (`[x.to('cpu')]`) as the reference; what changed is the
loop body.  The chain state lives on the GPU as one planar float32 tensor [2][B][H][W] (real plane,
imaginary plane), so that

  * the two score evaluations of a step (real part, imaginary part -- ALD_optimizers.py:227-228,
    440-441) are ONE batched forward of 2B images,
  * the Langevin update of both parts and the `L2Penalty` data-consistency step are ONE kernel
    (`ipdm_ald_sense_step`: z = x + step*g + sqrt(2 step)*n ;  x = z - kappa*(A^H A z - A^H y)),
  * a whole step (forward + update + schedule advance) is captured once in a CUDA graph and replayed
    for all L*n_steps_each steps; per-level scalars are read on the device from a schedule table.

Noise is drawn in-kernel -- Philox4x32-10 with key = seed and counter = (pixel, GLOBAL chain id, step): pass
`chain_ids` (one id per chain of the batch; default 0..B-1) and the same `seed` on every rank and chain i is the same
chain whether it runs alone, inside any batch, or on any rank of any world size -- unless `noise_fn(shape)` is given,
which injects host-chosen tensors in the reference's draw order (that is how the parity tests feed identical noise).
Without `seed` each call draws a fresh one from torch's global generator, like the reference's `torch.randn_like`
(so repeated calls give different chains and `torch.manual_seed` reproduces a run); a captured step graph reads the
seed from device memory and is not re-captured.
The reference's per-step prints (host syncs) and PNG snapshots are side effects, not results, and
are not reproduced.  Extra keyword arguments accepted by every `__call__`: `noise_fn`, `seed`, `chain_ids`,
`cuda_graph` (default True), `x_init` (override the initial state).
"""
import abc

import numpy as np
import torch
import torch.nn.functional as F

from . import ald_schedule
from .proximal_op import Proximal, L2Penalty, l2_kappa, _mask_kspace
from ..linear_transforms.undersampling_fourier import SENSE
from ... import _lib


def get_lh_weights(sigmas, start_time, curve_type="linear"):
    """Guidance-weight ramp (reference :23-38)."""
    assert 0 <= start_time <= 1
    lh_weights = torch.zeros_like(sigmas)
    if start_time == 1:
        return lh_weights
    start_idx = int(len(sigmas) * start_time)
    if curve_type == "linear":
        lh_weights[start_idx:] = torch.linspace(0, 1, len(sigmas) - start_idx, device=sigmas.device)
        return lh_weights
    raise NotImplementedError


def _default_device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _scalars(step, kappa=0.0, sigma=0.0, noise_on=True):
    step32 = torch.tensor(float(step), dtype=torch.float32)
    return _lib.AldScalars(float(step32), float(torch.sqrt(step32 * 2)) if noise_on else 0.0, float(kappa), float(sigma))


def _seed_and_chains(kwargs, n_chains, device, injected):
    """(seed, int32 device tensor of global chain ids) from the `seed` / `chain_ids` keyword arguments.  With
    injected noise no seed is needed and torch's global generator is left untouched (the injected draws come from it)."""
    seed = kwargs.pop("seed", None)
    seed = (0 if injected else _lib.fresh_seed()) if seed is None else int(seed)
    ids = kwargs.pop("chain_ids", None)
    ids = torch.arange(n_chains, dtype=torch.int32) if ids is None else torch.as_tensor(ids, dtype=torch.int32).reshape(-1)
    if ids.numel() != n_chains:
        raise ValueError(f"chain_ids has {ids.numel()} entries for {n_chains} chains")
    return seed, ids.to(device)


def _seed_tensor(seed, device):
    return torch.tensor([seed], dtype=torch.int64, device=device)


def data_transform(config, X):
    """Initial-state transform of the generic sampler (reference helpers/utils.py:212-226): dequantisation noise, the
    2x-1 rescale or the logit transform, image-mean subtraction.  The identity for every shipped config."""
    d = config.data
    if getattr(d, "uniform_dequantization", False):
        X = X / 256. * 255. + torch.rand_like(X) / 256.
    if getattr(d, "gaussian_dequantization", False):
        X = X + torch.randn_like(X) * 0.01
    if getattr(d, "rescaled", False):
        X = 2 * X - 1.
    elif getattr(d, "logit_transform", False):
        lam = 1e-6
        X = lam + (1 - 2 * lam) * X
        X = torch.log(X) - torch.log1p(-X)
    if hasattr(config, 'image_mean'):
        return X - config.image_mean.to(X.device)[None, ...]
    return X


def _to_planar(xc):
    """complex64 (B,1,H,W) -> planar float32 [2][B][H][W]"""
    xc = xc.to(torch.complex64).contiguous()
    B, C, H, W = xc.shape
    out = torch.empty((2, B * C, H, W), dtype=torch.float32, device=xc.device)
    _lib.check(_lib.lib().ipdm_c64_to_planar(xc.data_ptr(), out.data_ptr(), xc.numel(), _lib.stream()), "c64_to_planar")
    return out


def _to_complex(planar, shape):
    out = torch.empty(shape, dtype=torch.complex64, device=planar.device)
    _lib.check(_lib.lib().ipdm_planar_to_c64(planar.data_ptr(), out.data_ptr(), out.numel(), _lib.stream()), "planar_to_c64")
    return out


class _StepGraph:
    """One ALD step captured as a CUDA graph.  `prime()` runs `body()` once eagerly on a side stream
    (allocates the score net's buffers, builds tensor maps) and then captures it; it must be called
    while the persistent state tensors hold dummy data, because the warm-up execution mutates them.
    Afterwards every call is a pure replay."""

    def __init__(self, body):
        self.body, self.graph, self.launches = body, None, 0

    def prime(self):
        L = _lib.lib()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = L.ipdm_launch_count()                      # one-off work (weight packing) is behind us
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.body()
        self.launches = L.ipdm_launch_count() - before      # kernels of this library per replayed step
        return self

    def __call__(self):
        self.graph.replay()


class ALDOptimizer(abc.ABC):
    def __init__(self, x_mod_shape, scorenet, sigmas, params, config,
                 measurement=None, linear_tfm=None, clf=None, seg=None, device=None):
        """params: n_steps_each, step_lr, denoise, final_only   (reference :50-64)"""
        self.x_mod_shape = x_mod_shape
        self.scorenet = scorenet
        self.sigmas = sigmas
        self.params = params
        self.config = config
        self.measurement = measurement
        self.linear_tfm = linear_tfm
        self.clf = clf
        self.seg = seg
        self.device = device if device is not None else _default_device()
        self.launches_per_step = None
        self._fast_cache = {}

    # ---- hooks (same names as the reference) ----------------------------------------------------
    def preprocessing_steps(self, **kwargs):
        pass

    def init_x_mod(self):
        return torch.rand(*self.x_mod_shape).to(self.device)

    def init_estimation(self, x_mod, **kwargs):
        return x_mod

    def adjust_grad(self, grad, x_mod, **kwargs):
        return grad

    def _hooks_overridden(self):
        cls = type(self)
        return (cls.adjust_grad is not ALDOptimizer.adjust_grad or cls.init_estimation is not ALDOptimizer.init_estimation)

    def _range_check_on(self, x, labels, out):
        """Before a step graph is primed (its warm-up runs on dummy zeros): one eager score forward on the REAL initial state
        at the first -- largest -- noise level, so that the score network's first-forward range audit (operand exponent shift,
        DESIGN 2) looks at real activations and has settled before anything is captured."""
        self._score_into(x, labels, out)

    def _score_into(self, x, labels, out):
        if hasattr(self.scorenet, "forward_into"):
            return self.scorenet.forward_into(x, labels, out)
        out.copy_(self.scorenet(x, labels))
        return out

    # ---- generic loop (reference :66-137) ----------------------------------------------------------
    def __call__(self, **kwargs):
        torch.set_grad_enabled(False)
        noise_fn = kwargs.pop("noise_fn", None)
        use_graph = bool(kwargs.pop("cuda_graph", True))
        x_init = kwargs.pop("x_init", None)
        sigmas = self.sigmas
        n_steps_each = self.params["n_steps_each"]
        step_lr = self.params["step_lr"]
        L = _lib.lib()

        x_mod = self.init_x_mod() if x_init is None else x_init
        x_mod = data_transform(self.config, x_mod.to(self.device, torch.float32)).contiguous().clone()   # reference :86
        _lib.require_cuda(x_mod)
        B = x_mod.shape[0]
        seed, chain_ids = _seed_and_chains(kwargs, B, x_mod.device, noise_fn is not None)
        per_chain = x_mod[0].numel()
        self.preprocessing_steps(**kwargs)
        grad = torch.empty_like(x_mod)
        labels = torch.zeros(B, dtype=torch.long, device=x_mod.device)
        images = []
        hooks = self._hooks_overridden()
        fast = use_graph and noise_fn is None and not hooks and self.params["final_only"]
        if fast:
            sched_host = ald_schedule(sigmas, n_steps_each, step_lr)
            n_total = sched_host.shape[0]
            key = ("uncond", tuple(x_mod.shape), n_total, n_steps_each, x_mod.device)
            fc = self._fast_cache.get(key)
            if fc is None:
                self._range_check_on(x_mod, labels, torch.empty_like(x_mod))
                fc = {"x": torch.zeros_like(x_mod), "grad": torch.zeros_like(x_mod), "labels": torch.zeros_like(labels),
                      "sched": torch.zeros_like(sched_host, device=x_mod.device),
                      "cursor": torch.zeros(1, dtype=torch.int32, device=x_mod.device),
                      "seed": _seed_tensor(0, x_mod.device), "chain_ids": torch.zeros_like(chain_ids)}

                def body(fc=fc):
                    self._score_into(fc["x"], fc["labels"], fc["grad"])
                    _lib.check(L.ipdm_langevin_update(fc["x"].data_ptr(), fc["grad"].data_ptr(), None, None, fc["x"].numel(), None,
                                                      fc["sched"].data_ptr(), fc["cursor"].data_ptr(), None, 0,
                                                      _lib.rng(0, 0, fc["chain_ids"], per_chain, fc["seed"]), _lib.stream()), "langevin_update")
                    _lib.check(L.ipdm_ald_advance(fc["cursor"].data_ptr(), fc["labels"].data_ptr(), B, n_steps_each, _lib.stream()), "ald_advance")

                fc["step"] = _StepGraph(body).prime()
                self._fast_cache[key] = fc
            fc["seed"].fill_(seed)
            fc["chain_ids"].copy_(chain_ids)
            fc["x"].copy_(x_mod)
            fc["sched"].copy_(sched_host)
            fc["cursor"].zero_()
            fc["labels"].zero_()
            self.launches_per_step = fc["step"].launches
            for _ in range(n_total):
                fc["step"]()
            x_mod, grad, labels = fc["x"].clone(), fc["grad"], fc["labels"]
        else:
            k = 0
            for c, sigma in enumerate(sigmas):
                labels.fill_(c)
                step_size = step_lr * (sigma / sigmas[-1]) ** 2
                x_mod = self.init_estimation(x_mod, alpha=step_size, **kwargs)
                for s in range(n_steps_each):
                    self._score_into(x_mod, labels, grad)
                    g = self.adjust_grad(grad, x_mod, sigma=sigma, **kwargs)
                    noise = None if noise_fn is None else noise_fn(x_mod.shape).to(x_mod.device, torch.float32).contiguous()
                    _lib.check(L.ipdm_langevin_update(x_mod.data_ptr(), g.contiguous().data_ptr(), _lib.ptr(noise), None,
                                                      x_mod.numel(), _scalars(step_size), None, None, None, 0,
                                                      _lib.rng(seed, k, chain_ids, per_chain), _lib.stream()), "langevin_update")
                    k += 1
                    if not self.params["final_only"]:
                        images.append(x_mod.to('cpu'))
        if self.params["denoise"]:
            labels.fill_(len(sigmas) - 1)
            self._score_into(x_mod, labels, grad)
            _lib.check(L.ipdm_langevin_update(x_mod.data_ptr(), grad.data_ptr(), None, None, x_mod.numel(),
                                              _scalars(sigmas[-1] ** 2, noise_on=False), None, None, None, 0, None,
                                              _lib.stream()), "denoise")
            images.append(x_mod.to('cpu'))
        if self.params["final_only"]:
            return [x_mod.to('cpu')]
        return images


class ALDUnconditionalSampler(ALDOptimizer):
    pass


class _SenseChainMixin:
    """State + kernels shared by the two SENSE samplers."""

    def _fused_ok(self):
        return (isinstance(self.proximal, L2Penalty) and isinstance(self.linear_tfm, SENSE)
                and self.proximal.lin_tfm is self.linear_tfm)

    def _setup_sense(self, measurement5, device):
        """measurement5: (Nc, B', 1, H, W) complex64 on device.  Returns planar x0 = A^H y and b = A^H(mask*y)."""
        A = self.linear_tfm
        x0 = A.conj_op(measurement5)
        if self._fused_ok():
            b = A.conj_op_masked(_mask_kspace(A, measurement5.clone()))
        else:
            b = x0
        return _to_planar(x0), _to_planar(b)

    def _sense_step(self, state, grad, noise, bvec, scalars=None, sched=None, cursor=None, rng=None):
        A = self.linear_tfm
        mre, mim = A.device_maps(state.device)
        _, frames = A.device_mask(state.device)
        _, Bp, H, W = state.shape
        if frames not in (1, Bp):
            raise RuntimeError(f"mask has {frames} frames but the batch holds {Bp} images")
        plan = A.device_plan(state.device, H)
        _lib.check(_lib.lib().ipdm_ald_sense_step_plan(plan.handle, state.data_ptr(), grad.data_ptr(), _lib.ptr(noise), bvec.data_ptr(),
                                                       mre.data_ptr(), _lib.ptr(mim), mre.shape[0], Bp, H,
                                                       scalars, _lib.ptr(sched), _lib.ptr(cursor), rng, _lib.stream()),
                   "ald_sense_step")


class ALDInvSegProximalRealImag(_SenseChainMixin, ALDOptimizer):
    def __init__(self, proximal: Proximal, seg_start_time, seg_step_type, *args, **kwargs):
        super(ALDInvSegProximalRealImag, self).__init__(*args, **kwargs)
        self.proximal = proximal
        self.seg_start_time = seg_start_time
        self.seg_step_type = seg_step_type
        self.lh_weights = get_lh_weights(self.sigmas, self.seg_start_time, self.seg_step_type)
        self.if_print = False
        self.print_args = {}

    def adjust_grad(self, grad, m_mod, **kwargs):
        """grad + d/dx log p_seg(label | x) / sigma * seg_lamda   (reference :272-286); skipped when the
        weight is zero or no segmentation net is attached."""
        lamda = kwargs.get("seg_lamda", 0.)
        if self.seg is None or float(lamda) == 0.:
            return grad
        label, sigma, seg_mode = kwargs["label"], kwargs["sigma"], kwargs["seg_mode"]
        with torch.enable_grad():
            X = m_mod.detach().clone().requires_grad_(True)
            prob = torch.softmax(self.seg(X), dim=1)
            sel = torch.gather(prob, dim=1, index=label)
            torch.log(sel).sum().backward()
            g = X.grad
        if seg_mode == "FG":
            g = g * label
        return grad + g / sigma * lamda

    def post_processing(self, x_mod_real, x_mod_imag, **kwargs):
        """Generic (non-fused) data-consistency step on separate real / imaginary tensors (reference :288-327)."""
        x_mod = torch.complex(x_mod_real, x_mod_imag)
        coeff = kwargs["alpha"] * kwargs["lr_scaled"]
        x_mod = self.proximal(x_mod, self.measurement, coeff, 1.)
        return torch.real(x_mod), torch.imag(x_mod)

    def __call__(self, **kwargs):
        """kwargs: label, lamda, save_dir, lr_scaled, seg_mode  (+ noise_fn, seed, cuda_graph).
        `return_chain=True` (benchmarks) returns the primed fast-chain handle -- a dict whose "step" entry
        replays one captured ALD step on the loaded state -- instead of running the schedule."""
        torch.set_grad_enabled(False)
        noise_fn = kwargs.pop("noise_fn", None)
        use_graph = bool(kwargs.pop("cuda_graph", True))
        return_chain = bool(kwargs.pop("return_chain", False))
        sigmas = self.sigmas
        n_steps_each = self.params["n_steps_each"]
        step_lr = self.params["step_lr"]
        lr_scaled = kwargs.get("lr_scaled", 1.)
        L = _lib.lib()
        y = self.measurement.to(self.device)
        _lib.require_cuda(y)
        Nc, B, C, H, W = y.shape
        if C != 1:
            raise _lib.IpdmError("ALDInvSegProximalRealImag: C must be 1")
        seed, chain_ids = _seed_and_chains(kwargs, B, y.device, noise_fn is not None)
        plane_ids = torch.cat([2 * chain_ids, 2 * chain_ids + 1])      # the non-fused path updates the two planes as 2B samples
        state, bvec = self._setup_sense(y.to(torch.complex64).contiguous(), y.device)   # [2][B][H][W]
        self.preprocessing_steps(**kwargs)
        grad = torch.empty_like(state)
        labels = torch.zeros(2 * B, dtype=torch.long, device=state.device)
        x_flat = state.view(2 * B, 1, H, W)
        g_flat = grad.view(2 * B, 1, H, W)
        guided = self.seg is not None and bool((self.lh_weights != 0).any())
        hooked = type(self).adjust_grad is not ALDInvSegProximalRealImag.adjust_grad or \
            type(self).post_processing is not ALDInvSegProximalRealImag.post_processing
        fused = self._fused_ok() and not hooked
        fast = fused and use_graph and noise_fn is None and not guided
        if fast:
            kappa = l2_kappa(self.linear_tfm, state, step_lr * lr_scaled, 1.)
            sched_host = ald_schedule(sigmas, n_steps_each, step_lr, kappa)
            n_total = sched_host.shape[0]
            key = ("sense", B, H, W, n_total, n_steps_each, state.device)
            fc = self._fast_cache.get(key)
            if fc is None:
                self._range_check_on(x_flat, labels, g_flat)
                fc = {"state": torch.zeros_like(state), "grad": torch.zeros_like(state), "bvec": torch.zeros_like(state),
                      "labels": torch.zeros_like(labels), "sched": torch.zeros_like(sched_host, device=state.device),
                      "cursor": torch.zeros(1, dtype=torch.int32, device=state.device),
                      "seed": _seed_tensor(0, state.device), "chain_ids": torch.zeros_like(chain_ids)}

                def body(fc=fc):
                    self._score_into(fc["state"].view(2 * B, 1, H, W), fc["labels"], fc["grad"].view(2 * B, 1, H, W))
                    self._sense_step(fc["state"], fc["grad"], None, fc["bvec"], None, fc["sched"], fc["cursor"],
                                     _lib.rng(0, 0, fc["chain_ids"], 0, fc["seed"]))
                    _lib.check(L.ipdm_ald_advance(fc["cursor"].data_ptr(), fc["labels"].data_ptr(), 2 * B, n_steps_each, _lib.stream()), "ald_advance")

                fc["step"] = _StepGraph(body).prime()
                self._fast_cache[key] = fc
            fc["seed"].fill_(seed)
            fc["chain_ids"].copy_(chain_ids)
            fc["state"].copy_(state)
            fc["bvec"].copy_(bvec)
            fc["sched"].copy_(sched_host)
            fc["cursor"].zero_()
            fc["labels"].zero_()
            self.launches_per_step = fc["step"].launches
            self.fast_chain = fc
            if return_chain:
                return fc
            for _ in range(n_total):
                fc["step"]()
            state, grad, labels = fc["state"], fc["grad"], fc["labels"]
            x_flat, g_flat = state.view(2 * B, 1, H, W), grad.view(2 * B, 1, H, W)
        else:
            k = 0
            for c, sigma in enumerate(sigmas):
                labels.fill_(c)
                step_size = step_lr * (sigma / sigmas[-1]) ** 2
                w_seg = self.lh_weights[c]
                for s in range(n_steps_each):
                    self._score_into(x_flat, labels, g_flat)
                    if guided or hooked:
                        hk = dict(kwargs, sigma=sigma, seg_lamda=w_seg)
                        grad[0].copy_(self.adjust_grad(grad[0].unsqueeze(1), state[0].unsqueeze(1), **hk).squeeze(1))
                        grad[1].copy_(self.adjust_grad(grad[1].unsqueeze(1), state[1].unsqueeze(1), **hk).squeeze(1))
                    noise = None
                    if noise_fn is not None:  # reference draw order: real, then imaginary (:238-241)
                        nr = noise_fn((B, 1, H, W))
                        ni = noise_fn((B, 1, H, W))
                        noise = torch.stack([nr.reshape(B, H, W), ni.reshape(B, H, W)], 0).to(state.device, torch.float32).contiguous()
                    if fused:
                        kappa = l2_kappa(self.linear_tfm, state, step_lr * lr_scaled, 1.)
                        self._sense_step(state, grad, noise, bvec, _scalars(step_size, kappa, sigma), None, None,
                                         _lib.rng(seed, k, chain_ids))
                    else:
                        _lib.check(L.ipdm_langevin_update(state.data_ptr(), grad.data_ptr(), _lib.ptr(noise), None, state.numel(),
                                                          _scalars(step_size), None, None, None, 0,
                                                          _lib.rng(seed, k, plane_ids, H * W), _lib.stream()), "langevin_update")
                        xr, xi = self.post_processing(state[0].unsqueeze(1), state[1].unsqueeze(1), alpha=step_lr, sigma=sigma, **kwargs)
                        state[0].copy_(xr.squeeze(1))
                        state[1].copy_(xi.squeeze(1))
                    k += 1
        if self.params["denoise"]:
            labels.fill_(len(sigmas) - 1)
            self._score_into(x_flat, labels, g_flat)
            _lib.check(L.ipdm_langevin_update(state.data_ptr(), grad.data_ptr(), None, None, state.numel(),
                                              _scalars(sigmas[-1] ** 2, noise_on=False), None, None, None, 0, None,
                                              _lib.stream()), "denoise")
        x_mod = _to_complex(state, (B, 1, H, W))
        self.final_state = x_mod
        return [x_mod.to('cpu')]


class ALD2DTime(_SenseChainMixin, ALDOptimizer):
    def __init__(self, proximal: Proximal, scorenet_T, sigmas_T, *args, **kwargs):
        """x_mod_shape: (B, T, C, H, W); measurement: (num_sens, B, T, C, H, W)   (reference :330-349)"""
        super(ALD2DTime, self).__init__(*args, **kwargs)
        self.proximal = proximal
        self.scorenet_T = scorenet_T
        self.sigmas_T_orig = sigmas_T
        # temporal schedule nearest-interpolated onto the tail of the spatial one, -1 = "skip" (quirk Q14)
        n = int((self.sigmas <= sigmas_T[0]).sum())
        self.sigmas_T = torch.ones_like(self.sigmas) * (-1)
        if n > 0:
            self.sigmas_T[-n:] = F.interpolate(sigmas_T.view(1, 1, -1).float(), n, mode="nearest").squeeze().to(self.sigmas.device)
        if self.scorenet_T is not None:
            self.scorenet_T.sigmas = self.sigmas_T
            self.win_size = int(np.sqrt(self.scorenet_T.config.data.channels))
        self.finite_diff = None

    def __call__(self, **kwargs):
        """kwargs: save_dir, lr_scaled, mode_T in {"none", "tv", "tv-only", "diffusion1d", "diffusion1d-only"}, lamda_T,
        if_random_shift (+ noise_fn, seed, cuda_graph).  "diffusion1d" = the learned temporal prior `scorenet_T`
        (NCSN3DShallow on k x k x T patches, reference :463-502)."""
        torch.set_grad_enabled(False)
        noise_fn = kwargs.pop("noise_fn", None)
        use_graph = bool(kwargs.pop("cuda_graph", True))
        mode_T = kwargs.get("mode_T", "diffusion1d")
        lamda_T = float(kwargs.get("lamda_T", 1.))
        lr_scaled = kwargs["lr_scaled"]
        diffusion = "diffusion1d" in mode_T
        if diffusion and self.scorenet_T is None:
            raise _lib.IpdmError("mode_T='diffusion1d' needs scorenet_T (NCSN3DShallow)")
        random_shift = bool(kwargs.get("if_random_shift", False))
        skip_spatial = mode_T in ("tv-only", "diffusion1d-only")
        if skip_spatial:
            self.sigmas_T = self.sigmas_T_orig
            self.sigmas = self.sigmas_T_orig
            if self.scorenet_T is not None:
                self.scorenet_T.sigmas = self.sigmas_T_orig
        sigmas = self.sigmas
        n_steps_each = self.params["n_steps_each"]
        step_lr = self.params["step_lr"]
        L = _lib.lib()
        self.preprocessing_steps(**kwargs)
        y6 = self.measurement.to(self.device).to(torch.complex64)
        _lib.require_cuda(y6)
        Nc, B, T, C, H, W = y6.shape
        if C != 1:
            raise _lib.IpdmError("ALD2DTime: C must be 1")
        if not self._fused_ok():
            raise _lib.IpdmError("ALD2DTime: only L2Penalty over SENSE is implemented")
        y = y6.reshape(Nc, B * T, C, H, W).contiguous()
        seed, chain_ids = _seed_and_chains(kwargs, B, y.device, noise_fn is not None)
        state, bvec = self._setup_sense(y, y.device)                      # [2][B*T][H][W]
        BT = B * T
        # noise streams: frame t of chain i is sample i*T + t; the generic update sees the two planes as 2*B*T samples
        frame_ids = (chain_ids.view(B, 1) * T + torch.arange(T, dtype=torch.int32, device=y.device).view(1, T)).reshape(-1).contiguous()
        plane_ids = torch.cat([2 * frame_ids, 2 * frame_ids + 1])
        grad = torch.zeros_like(state)
        ksz = self.win_size if diffusion else 1
        P2 = 2 * B * (H // ksz) * (W // ksz) if diffusion else 0      # temporal-prior patches (real and imaginary planes)
        labels_all = torch.zeros(max(2 * BT, P2), dtype=torch.long, device=state.device)
        labels = labels_all[:2 * BT]
        x_flat, g_flat = state.view(2 * BT, 1, H, W), grad.view(2 * BT, 1, H, W)
        kappa = l2_kappa(self.linear_tfm, state, step_lr * lr_scaled, 1.)
        tv = "tv" in mode_T
        prox_only = _lib.AldScalars(0.0, 0.0, float(kappa), 0.0)
        fast = use_graph and noise_fn is None
        sig_T = self.sigmas_T.detach().float().cpu()
        temporal_on = [diffusion and float(sig_T[c]) != -1.0 for c in range(len(sigmas))]
        SEED_T = 0x5bd1e995                                            # the temporal prior's stream: seed ^ SEED_T
        if diffusion:
            vol = torch.zeros(P2, ksz, T, ksz, dtype=torch.float32, device=state.device)
            gvol = torch.zeros_like(vol)
            npp = (H // ksz) * (W // ksz)                               # patches per plane and chain, volumes ordered (plane, chain, patch)
            vol_ids = ((2 * chain_ids.view(1, B, 1) + torch.arange(2, dtype=torch.int32, device=y.device).view(2, 1, 1)) * npp
                       + torch.arange(npp, dtype=torch.int32, device=y.device).view(1, 1, npp)).reshape(-1).contiguous()
        ids = {"frame": frame_ids, "plane": plane_ids, "vol": vol_ids if diffusion else None, "seed_dev": None}

        def rng_for(kind, k, temporal=False):
            """ipdm_rng of one launch: eager steps pass the seed by value, captured steps read it from device memory"""
            elems = {"frame": 0, "plane": H * W, "vol": ksz * T * ksz}[kind]
            base = SEED_T if temporal else 0
            if ids["seed_dev"] is not None:
                return _lib.rng(base, 0, ids[kind], elems, ids["seed_dev"])
            return _lib.rng(seed ^ base, k, ids[kind], elems)

        def fold(unfold, sh, sw, shifts, cursor):
            if shifts is not None:      # captured step: this step's roll comes from the device table
                _lib.check(L.ipdm_patch_fold_sched(state.data_ptr(), vol.data_ptr(), B, T, H, W, ksz, shifts.data_ptr(), cursor.data_ptr(),
                                                   unfold, _lib.stream()), "patch_fold_sched")
            else:
                _lib.check(L.ipdm_patch_fold(state.data_ptr(), vol.data_ptr(), B, T, H, W, ksz, sh, sw, unfold, _lib.stream()), "patch_fold")

        def temporal_diffusion(c, k, sched_T=None, cursor=None, lab=None, shifts=None):
            """fold -> scorenet_T on real and imaginary patches -> Langevin update of the patches -> unfold  (:463-502)"""
            sh, sw = (tuple(np.random.randint(0, ksz, (2,)).tolist()) if random_shift and shifts is None else (0, 0))
            fold(0, sh, sw, shifts, cursor)
            lab = labels_all[:P2] if lab is None else lab
            if hasattr(self.scorenet_T, "forward_into"):
                self.scorenet_T.forward_into(vol, lab, gvol)
            else:
                flat = vol.permute(0, 1, 3, 2).reshape(P2, ksz * ksz, T)
                gvol.copy_(self.scorenet_T(flat, lab).reshape(P2, ksz, ksz, T).permute(0, 1, 3, 2))
            if sched_T is not None:
                _lib.check(L.ipdm_langevin_update(vol.data_ptr(), gvol.data_ptr(), None, None, vol.numel(), None, sched_T.data_ptr(),
                                                  cursor.data_ptr(), None, 0, rng_for("vol", 0, True), _lib.stream()), "langevin_update_T")
            else:
                step_T = step_lr * (self.sigmas_T[c].float().cpu() / sig_T[-1]) ** 2 * lamda_T
                nz = None
                if noise_fn is not None:   # reference draws (B', kx*ky, T) real then imaginary (:484-485)
                    nr, ni = noise_fn((P2 // 2, ksz * ksz, T)), noise_fn((P2 // 2, ksz * ksz, T))
                    nz = torch.cat([nr, ni]).reshape(P2, ksz, ksz, T).permute(0, 1, 3, 2).to(state.device, torch.float32).contiguous()
                _lib.check(L.ipdm_langevin_update(vol.data_ptr(), gvol.data_ptr(), _lib.ptr(nz), None, vol.numel(), _scalars(step_T),
                                                  None, None, None, 0, rng_for("vol", k, True), _lib.stream()), "langevin_update_T")
            fold(1, sh, sw, shifts, cursor)

        def one_step(c, k, noise, sched=None, cursor=None, sched_T=None, with_T=None, shifts=None):
            """spatial_step (:428-449) -> temporal_step (:452-502) -> proximal_step (:543-554)"""
            if not skip_spatial:
                self._score_into(x_flat, labels, g_flat)
            if not tv and not diffusion:
                # Langevin update and L2-penalty step in one kernel
                if sched is not None:
                    self._sense_step(state, grad, None, bvec, None, sched, cursor, rng_for("frame", 0))
                else:
                    step_size = step_lr * (sigmas[c] / sigmas[-1]) ** 2
                    self._sense_step(state, grad, noise, bvec, _scalars(step_size, kappa, sigmas[c]), None, None, rng_for("frame", k))
                return
            if not skip_spatial:
                if sched is not None:
                    _lib.check(L.ipdm_langevin_update(state.data_ptr(), grad.data_ptr(), None, None, state.numel(), None,
                                                      sched.data_ptr(), cursor.data_ptr(), None, 0, rng_for("plane", 0), _lib.stream()), "langevin_update")
                else:
                    step_size = step_lr * (sigmas[c] / sigmas[-1]) ** 2
                    _lib.check(L.ipdm_langevin_update(state.data_ptr(), grad.data_ptr(), _lib.ptr(noise), None, state.numel(),
                                                      _scalars(step_size), None, None, None, 0, rng_for("plane", k), _lib.stream()), "langevin_update")
            if tv:
                _lib.check(L.ipdm_temporal_tv_step(state.data_ptr(), B, T, H * W, lamda_T, _lib.stream()), "temporal_tv_step")
            elif temporal_on[c] if with_T is None else with_T:
                temporal_diffusion(c, k, sched_T, cursor, shifts=shifts)
            # data consistency only: step = noise_scale = 0 turns the fused kernel into x - kappa*(A^H A x - b)
            self._sense_step(state, grad, None, bvec, prox_only, None, None, None)

        if fast and not skip_spatial:
            sched_host = ald_schedule(sigmas, n_steps_each, step_lr, kappa)
            n_total = sched_host.shape[0]
            key = ("cine", B, T, H, W, n_total, n_steps_each, mode_T, lamda_T, float(kappa), bool(diffusion and random_shift), state.device)
            fc = self._fast_cache.get(key)
            real = (state, bvec)
            if fc is None:
                self._range_check_on(x_flat, labels, g_flat)      # the spatial prior on the real initial frames
                fc = {"state": torch.zeros_like(state), "grad": torch.zeros_like(state), "bvec": torch.zeros_like(state),
                      "labels": torch.zeros_like(labels_all), "sched": torch.zeros_like(sched_host, device=state.device),
                      "cursor": torch.zeros(1, dtype=torch.int32, device=state.device),
                      "seed": _seed_tensor(0, state.device), "frame": torch.zeros_like(frame_ids), "plane": torch.zeros_like(plane_ids)}
                if diffusion:
                    fc["vol_ids"] = torch.zeros_like(vol_ids)
                if diffusion:
                    fc.update(vol=vol, gvol=gvol, sched_T=torch.zeros_like(sched_host, device=state.device))
                    if random_shift:
                        fc["shifts"] = torch.zeros(n_total, 2, dtype=torch.int32, device=state.device)
                self._fast_cache[key] = fc
            state, grad, bvec, labels_all = fc["state"], fc["grad"], fc["bvec"], fc["labels"]
            fc["seed"].fill_(seed)
            fc["frame"].copy_(frame_ids)
            fc["plane"].copy_(plane_ids)
            if diffusion:
                fc["vol_ids"].copy_(vol_ids)
            ids.update(frame=fc["frame"], plane=fc["plane"], vol=fc.get("vol_ids"), seed_dev=fc["seed"])
            labels = labels_all[:2 * BT]
            if diffusion:
                vol, gvol = fc["vol"], fc["gvol"]
            x_flat, g_flat = state.view(2 * BT, 1, H, W), grad.view(2 * BT, 1, H, W)
            if "step" not in fc:
                def body(with_T, fc=fc):
                    one_step(0, 0, None, fc["sched"], fc["cursor"], fc.get("sched_T"), with_T, fc.get("shifts"))
                    _lib.check(L.ipdm_ald_advance(fc["cursor"].data_ptr(), fc["labels"].data_ptr(), fc["labels"].numel(), n_steps_each,
                                                  _lib.stream()), "ald_advance")
                # the temporal prior is skipped on the levels whose remapped sigma_T is -1 (Q14): one captured step without
                # it, one with it; the host picks per level
                fc["step"] = _StepGraph(lambda: body(False)).prime()
                if any(temporal_on):
                    fc["step_T"] = _StepGraph(lambda: body(True)).prime()
            state.copy_(real[0])
            bvec.copy_(real[1])
            fc["sched"].copy_(sched_host)
            if diffusion:
                fc["sched_T"].copy_(ald_schedule(torch.where(sig_T > 0, sig_T, sig_T[-1]), n_steps_each, step_lr * lamda_T))
            if "shifts" in fc:
                # the rolls of the whole chain, drawn in the order the per-step path (and the reference, :466-470) draws them:
                # one np.random.randint(0, k, (2,)) per step that runs the temporal prior
                tab = np.zeros((n_total, 2), dtype=np.int32)
                for i in range(n_total):
                    if temporal_on[i // n_steps_each]:
                        tab[i] = np.random.randint(0, ksz, (2,))
                fc["shifts"].copy_(torch.from_numpy(tab))
            fc["cursor"].zero_()
            labels_all.zero_()
            self.launches_per_step = fc["step_T"].launches if "step_T" in fc else fc["step"].launches
            for i in range(n_total):
                fc["step_T" if temporal_on[i // n_steps_each] else "step"]()
        else:
            k = 0
            for c in range(len(sigmas)):
                labels_all.fill_(c)
                for s in range(n_steps_each):
                    noise = None
                    if noise_fn is not None and not skip_spatial:  # both noises are drawn before either update (:442-443)
                        nr = noise_fn((BT, 1, H, W))
                        ni = noise_fn((BT, 1, H, W))
                        noise = torch.stack([nr.reshape(BT, H, W), ni.reshape(BT, H, W)], 0).to(state.device, torch.float32).contiguous()
                    one_step(c, k, noise)
                    k += 1
        x_mod = _to_complex(state, (B, T, C, H, W))
        self.final_state = x_mod
        return [x_mod.to("cpu")]
