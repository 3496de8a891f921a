"""Parity cases shared by the CPU host-logic suite (C ABI emulated, tests/emu_lib.py) and the GPU
suite (real libipdm_b200.so).  Every case compares the PRODUCT package against the golden fixtures
generated from the reference (tests/golden/) and/or the oracle on the same seeded inputs.

Tolerances (relative L2): fp32 FFT/SENSE/prox work 1e-5 (north_star: SENSE fwd/adj within 1e-5);
anything that goes through the score network uses f16 tensor-core operands with fp32 accumulation:
score itself 5e-3, one ALD step / short chains 1e-4 on x (north_star: single ALD step within 1e-4).
"""
import json
import os
import types

import numpy as np
import torch

from conftest import rel_l2, GOLDEN
from oracle import mri_ops as M
from oracle import ald as OALD
from oracle import scorenet as SN
from oracle.fixture_inputs import crandn, rrand, rrandn, phantom

import inverseproblemwithdiffusionmodel_b200 as P
from inverseproblemwithdiffusionmodel_b200.ncsn.linear_transforms import i2k_complex, k2i_complex, generate_mask
from inverseproblemwithdiffusionmodel_b200.ncsn.linear_transforms.undersampling_fourier import (
    SENSE, RandomUndersamplingFourier, keep_center_mask)
from inverseproblemwithdiffusionmodel_b200.ncsn.linear_transforms.finite_diff import FiniteDiff
from inverseproblemwithdiffusionmodel_b200.ncsn.models import get_sigmas, anneal_Langevin_dynamics
from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsnv2 import NCSNv2, NCSNv2Deeper, NCSNv2Deepest
from inverseproblemwithdiffusionmodel_b200.ncsn.models.proximal_op import L2Penalty, SingleCoil, Constrained, get_proximal
from inverseproblemwithdiffusionmodel_b200.ncsn.models import ALD_optimizers as ALD
from inverseproblemwithdiffusionmodel_b200.sde.sampling import AnnealedLangevinDynamics, LangevinCorrector
from inverseproblemwithdiffusionmodel_b200.ncsn.models import MAP_optimizers as MAP

TOL32 = 1e-5
# Score of a whole network vs the fp32 oracle / reference.  f16 convolution operands cost ~8e-4 relative L2, the 16-bit
# residual stream of the tensor-core path takes it to ~1.3e-3 (measured on the GPU: 1.37e-3 Deepest, 32^2; oracle emulation
# on the CPU: 1.25e-3).  NB an emulation of the kernels' rounding points cannot be matched more tightly than that: in an
# f16-operand network a 1e-7 relative perturbation of the convolution results (summation order) flips roundings, the
# flips perturb the next layer by more, and after a few layers two runs differ by the full rounding-noise level (measured
# with the oracle: 1e-7 in -> 7.6e-4 out; in fp32 arithmetic the same perturbation gives 1e-6).  So the emulation bound is the
# noise level times ~sqrt(2), and regressions of the epilogues are caught by the per-kernel tests, which are bit-level.
TOL_SCORE = 2e-3
TOL_SCORE_EMU = 2.2e-3   # vs the oracle with f16-rounded operands (+ 16-bit stream): two realisations of the same rounding noise
TOL_X = 1e-4


def G(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def specs():
    with open(os.path.join(GOLDEN, "state_dict_specs.json")) as f:
        return json.load(f)


def ns(**kw):
    return types.SimpleNamespace(**kw)


def make_config(dataset, ngf, image_size, num_classes, sigma_begin, sigma_end=0.01, device="cpu"):
    """The nested-Namespace config the reference reads (ncsn/configs/*.yml), reduced to the used keys."""
    model = ns(sigma_begin=sigma_begin, num_classes=num_classes, sigma_end=sigma_end, sigma_dist="geometric",
               normalization="InstanceNorm++", nonlinearity="elu", ngf=ngf, ema=True, ema_rate=0.999, spec_norm=False)
    data = ns(dataset=dataset, image_size=image_size, channels=1, logit_transform=False, uniform_dequantization=False,
              gaussian_dequantization=False, random_flip=True, rescaled=False)
    recons = ns(sigma_dist="geometric", sigma_begin=sigma_begin, num_classes=num_classes, sigma_end=sigma_end)
    return ns(model=model, data=data, recons=recons, device=torch.device(device))


def build_net(cls, spec_name, seed, cfg, dev):
    net = cls(cfg)
    spec = [(k, tuple(s)) for k, s in specs()[spec_name]]
    assert [k for k, _ in spec] == list(net.state_dict().keys())
    Pd = SN.synth_state_dict(spec, seed, net.sigmas.cpu())
    net.load_state_dict(Pd)
    return net.to(dev).eval(), Pd


# ------------------------------------------------------------------------------------------------ FFT / SENSE
def case_fft(dev):
    g = G("linear_ops")
    for n in (16, 32):
        x = crandn(1100 + n, 2, 1, n, n).to(dev)
        assert rel_l2(i2k_complex(x).cpu(), g[f"fft_i2k_{n}"]) < TOL32
        assert rel_l2(k2i_complex(x).cpu(), g[f"fft_k2i_{n}"]) < TOL32
    x = rrandn(1199, 1, 1, 8, 32).to(dev)
    assert rel_l2(i2k_complex(x).cpu(), g["fft_i2k_rect"]) < TOL32
    assert rel_l2(k2i_complex(x).cpu(), g["fft_k2i_rect"]) < TOL32


def case_setup_bit_exact():
    g = G("linear_ops")
    assert np.array_equal(generate_mask(24, 128, sw=0.196, sm=0.5, sa=0.02, seed=3).numpy(), g["mask_R8_T24_N128_seed3"])
    assert np.array_equal(generate_mask(1, 64, seed=5).numpy(), g["mask_T1_N64_seed5"])
    for W in (32, 128, 256):
        F1 = RandomUndersamplingFourier(40, 1 / 64, (1, W, W), seed=0)
        assert F1.mask.dtype == torch.bool and tuple(F1.mask.shape) == (24, 1, 1, W)
        assert np.array_equal(F1.mask.numpy(), g[f"mask_live_W{W}_seed0"])
    A = SENSE("exp", 4, 40, 1 / 64, (1, 32, 32), 0)
    assert A.sens_maps.dtype == torch.float64
    assert np.allclose(A.sens_maps.numpy(), g["coil_maps_32_seed0"], rtol=0, atol=1e-14)
    assert np.array_equal(keep_center_mask(32, 4, 1 / 8, 0).numpy(), G("sense_prox")["kc_mask"])
    s = get_sigmas(make_config("ACDC", 128, 256, 2311, 348))
    assert torch.equal(s, OALD.geometric_sigmas(348, 0.01, 2311))


def case_sense(dev):
    g = G("sense_prox")
    n = 32
    A2 = SENSE("exp", 2, 40, 1 / 64, (1, n, n), 7)
    x24 = crandn(1201, 24, 1, n, n).to(dev)
    S = A2(x24)
    assert tuple(S.shape) == (2, 24, 1, n, n) and S.dtype == torch.complex64
    assert rel_l2(S[:, [0, 5, 23]].cpu(), g["live_S_frames_0_5_23"]) < TOL32
    assert rel_l2(A2.conj_op(S).cpu(), g["live_adj"]) < TOL32
    assert rel_l2(A2.conj_op_masked(S).cpu(), g["live_adj"]) < TOL32
    assert rel_l2(A2.SSOS(S).cpu(), g["live_ssos"]) < TOL32
    # Q2: a single image against the 24-frame mask silently becomes 24 undersamplings
    S1 = A2(x24[:1])
    assert tuple(S1.shape) == (2, 24, 1, n, n)
    try:
        A2(x24[:2])
        raise AssertionError("batch 2 against a 24-frame mask must fail like the reference broadcast")
    except RuntimeError:
        pass
    A = SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
    assert rel_l2(A.conj_op(crandn(1202, 4, 2, 1, n, n).to(dev)).cpu(), g["dense_adj"]) < TOL32   # Q3: no mask
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    x3 = crandn(1203, 3, 1, n, n).to(dev)
    S3 = A(x3)
    assert rel_l2(S3.cpu(), g["kc_S"]) < TOL32
    assert rel_l2(A.conj_op(S3).cpu(), g["kc_adj"]) < TOL32
    assert rel_l2(A.log_lh_grad(x3, S3 * 0.5, 0.7).cpu(), g["kc_loglh"]) < TOL32
    return A, x3, S3


def case_prox(dev):
    g = G("sense_prox")
    n = 32
    A, x3, S3 = case_sense(dev)
    z = crandn(1204, 3, 1, n, n).to(dev)
    prox = get_proximal("L2Penalty")(A)
    for tag, alpha in (("a1", 1.0), ("a1e3", 1e3)):
        ref = torch.as_tensor(g[f"l2_{tag}"])
        out = prox(z, S3, alpha, 1.0).cpu()
        assert rel_l2(out, ref) < TOL32
        assert rel_l2(out - z.cpu(), ref - z.cpu()) < 1e-4      # the data-consistency move itself
    F1 = RandomUndersamplingFourier(4, 1 / 8, (1, n, n), seed=0)
    F1.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    S1 = F1(x3)
    assert rel_l2(S1.cpu(), g["sc_S"]) < TOL32
    xs = SingleCoil(F1)(z, S1, 0.8, 1.0)
    assert rel_l2(xs.cpu(), g["sc_prox"]) < TOL32
    assert rel_l2(L2Penalty(F1)(z, S1, 2.0, 1.0).cpu(), g["sc_l2"]) < TOL32
    assert rel_l2(Constrained(F1)(z, S1, 0.3).cpu(), g["sc_proj"]) < TOL32
    assert float(SingleCoil(F1).check_solution(xs, z, S1, 0.8, 1.0)) < 1e-7
    # two SGD steps against the oracle's autograd implementation
    fwd = lambda v: M.sense_forward(v, A.sens_maps, A.random_under_fourier.mask)
    ref2 = M.l2_prox_sgd(fwd, z.cpu(), S3.cpu(), 50.0, 1.0, num_steps=2)
    assert rel_l2(prox(z, S3, 50.0, 1.0, num_steps=2).cpu(), ref2) < TOL32


def case_tv(dev):
    x = rrandn(31, 2, 6, 1, 8, 8).to(dev)
    g = FiniteDiff(dims=1).log_lh_grad(x, lamda=0.3)
    assert rel_l2(g.cpu(), OALD.temporal_tv_grad(x.cpu(), 0.3)) < 1e-6


# ------------------------------------------------------------------------------------------------ score network
def case_scorenet_small(dev):
    g = G("scorenet")
    cfg = make_config("ACDC", 8, 32, 12, 30.0)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 1, cfg, dev)
    out = net((rrand(1301, 2, 1, 32, 32) * 3 - 1).to(dev), torch.tensor([0, 7]).to(dev))
    assert rel_l2(out.cpu(), g["deepest_out"]) < TOL_SCORE
    cfg = make_config("MNIST", 8, 28, 12, 30.0)
    net, _ = build_net(NCSNv2, "NCSNv2_ngf8_28", 2, cfg, dev)
    out = net(rrand(1302, 2, 1, 28, 28).to(dev), torch.tensor([11, 3]).to(dev))
    assert rel_l2(out.cpu(), g["v2_out"]) < TOL_SCORE
    # a non-contiguous real view of a complex tensor is a legal input (ALD_optimizers.py:191,227)
    xc = crandn(5, 2, 1, 28, 28).to(dev)
    a = net(torch.real(xc), torch.tensor([1, 1]).to(dev))
    b = net(torch.real(xc).contiguous(), torch.tensor([1, 1]).to(dev))
    assert torch.equal(a, b)


def case_operand_shift(dev):
    """Operand exponent shift: (a) forced by IPDM_OPERAND_SHIFT it changes nothing beyond f16 rounding of shifted
    values -- same golden score; (b) an input whose activations leave the f16 range makes the first forward escalate by itself
    (warning), and the result then agrees with the fp32 oracle, while the clipped run (IPDM_ALLOW_F16_SATURATION) does not."""
    import os, warnings
    g = G("scorenet")
    cfg = make_config("ACDC", 8, 32, 12, 30.0)
    x, y = (rrand(1301, 2, 1, 32, 32) * 3 - 1).to(dev), torch.tensor([0, 7]).to(dev)
    os.environ["IPDM_OPERAND_SHIFT"] = "6"
    try:
        net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 1, cfg, dev)
        out = net(x, y)
        plan = next(iter(net._plans.values()))
        assert plan.shift == 6 and not plan.t16
    finally:
        del os.environ["IPDM_OPERAND_SHIFT"]
    assert rel_l2(out.cpu(), g["deepest_out"]) < TOL_SCORE
    # (b) |x| ~ 3e5: the un-normalised decoder leaves the f16 range
    big = (rrandn(78, 2, 1, 32, 32) * 3e5)
    net, Pd = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 1, cfg, dev)
    with torch.no_grad():
        ref = SN.score_forward("NCSNv2Deepest", Pd, big, y.cpu())
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        out = net(big.to(dev), y)
    plan = next(iter(net._plans.values()))
    assert plan.shift >= 6 and any("operand exponent shift" in str(w.message) for w in wlist), plan.shift
    e_shift = rel_l2(out.cpu(), ref)
    os.environ["IPDM_ALLOW_F16_SATURATION"] = "1"
    try:
        net2, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 1, cfg, dev)
        clipped = net2(big.to(dev), y)
    finally:
        del os.environ["IPDM_ALLOW_F16_SATURATION"]
    e_clip = rel_l2(clipped.cpu(), ref)
    print(f"operand shift {plan.shift}: rel-L2 vs fp32 oracle {e_shift:.2e}; clipped run {e_clip:.2e}")
    assert bool(torch.isfinite(out).all())
    if str(dev) != "cpu":      # the kernels saturate (finite); the CPU emulator's casts go to inf
        assert bool(torch.isfinite(clipped).all())
    assert e_shift < 5e-3 and not (e_clip <= 10 * e_shift), (e_shift, e_clip)


def case_ncsn3d_shallow(dev):
    """NCSN3DShallow (the learned temporal prior) at full width against the reference's output:
    5-D and flattened inputs, state-dict layout, and the f16-operand emulation of the oracle for the tight bound."""
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
    g = G("ncsn3d")
    cfg = make_config("CINE127", 128, 24, 12, 40.0)
    cfg.data.channels, cfg.data.channels_3d = 64, 1
    net, Pd = build_net(NCSN3DShallow, "NCSN3DShallow_ngf128", 12, cfg, dev)
    x, y = rrand(1701, 2, 1, 8, 8, 24), torch.tensor([2, 9])
    out = net(x.to(dev), y.to(dev))
    assert out.shape == (2, 1, 8, 8, 24)
    e_ref = rel_l2(out.cpu(), g["shallow_out"])
    assert e_ref < TOL_SCORE, e_ref
    flat = net(x.reshape(2, 64, 24).to(dev), y.to(dev))
    assert flat.shape == (2, 64, 24) and torch.equal(flat.reshape(2, 1, 8, 8, 24), out)
    SN.OPERAND_ROUND = torch.float16
    try:
        with torch.no_grad():
            emu = SN.score_forward_3d_shallow(Pd, x, y)
    finally:
        SN.OPERAND_ROUND = None
    e_emu = rel_l2(out.cpu(), emu)
    print(f"NCSN3DShallow: rel-L2 vs reference fp32 {e_ref:.2e}, vs f16-operand emulation {e_emu:.2e}")
    assert e_emu < 2e-3, e_emu                     # same operand rounding; rounding-boundary flips and accumulation order remain


def case_sampler_cine_diffusion(dev):
    """ALD2DTime with the learned temporal prior (mode_T='diffusion1d') against chains of the
    unmodified reference: injected torch noise in the reference's draw order, np.random frame rolls, the Q14 sigma remap
    (six spatial levels skip the temporal step, four run it)."""
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
    g = G("ncsn3d")
    n, T = 32, 8
    cfg = make_config("CINE127", 8, n, 10, 20.0, device=dev)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 5, cfg, dev)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1702, T, 1, n, n).to(dev)).reshape(4, 1, T, 1, n, n)
    cfg_T = make_config("CINE127", 128, T, 6, 0.2, device=dev)
    cfg_T.data.channels, cfg_T.data.channels_3d = 64, 1
    sig_T = get_sigmas(cfg_T)
    params = {"n_steps_each": 1, "step_lr": 1e-4}
    draw = lambda shape: torch.randn(*shape)
    for tag, shift in (("fixed", False), ("shift", True)):
        net_T, _ = build_net(NCSN3DShallow, "NCSN3DShallow_ngf128", 13, cfg_T, dev)
        sampler = ALD.ALD2DTime(L2Penalty(A), net_T, sig_T, (1, T, 1, n, n), net, sig, params, cfg,
                                measurement=meas, linear_tfm=A, device=torch.device(dev))
        assert tuple(net_T.sigmas.shape) == tuple(sig.shape) and int((net_T.sigmas == -1).sum()) == 6
        torch.manual_seed(304)
        np.random.seed(11)
        res = sampler(save_dir="/tmp", lr_scaled=1e4, mode_T="diffusion1d", lamda_T=0.5, if_random_shift=shift, noise_fn=draw)
        err = rel_l2(res[0], g[f"cine_diffusion_{tag}"])
        assert err < 1e-3, (tag, err)                       # same tolerance and reasoning as case_sampler_cine
    # MAP baseline with the learned temporal prior (MAPOptimizer2DTime mode_T='diffusion1d', reference :284-306)
    net_T, _ = build_net(NCSN3DShallow, "NCSN3DShallow_ngf128", 13, cfg_T, dev)
    x0 = A.conj_op(meas.reshape(4, T, 1, n, n)).reshape(1, T, 1, n, n).clone()
    mp = dict(lr=5e-3, opt_class=torch.optim.Adam, num_iters=2, num_plot_times=1, win_size=8, prior_weight=1.0,
              spatial_step_weight=0.7, temporal_step_weight=0.3, save_dir="/tmp", opt_params={"betas": (0.5, 0.5)},
              mode_T="diffusion1d", if_random_shift=False)
    rec = MAP.MAPOptimizer2DTime(x0, meas, net, net_T, A, None, mp)()
    assert rel_l2(rec, g["map2dt_diffusion"]) < 1e-3
    # in-kernel Philox noise + captured step graphs (one with, one without the temporal step): finite, and close to the
    # injected-noise chain in distribution (same start, same schedule)
    net_T, _ = build_net(NCSN3DShallow, "NCSN3DShallow_ngf128", 13, cfg_T, dev)
    sampler = ALD.ALD2DTime(L2Penalty(A), net_T, sig_T, (1, T, 1, n, n), net, sig, params, cfg,
                            measurement=meas, linear_tfm=A, device=torch.device(dev))
    fast = sampler(save_dir="/tmp", lr_scaled=1e4, mode_T="diffusion1d", lamda_T=0.5, if_random_shift=False, seed=3)[0]
    assert torch.isfinite(fast.abs()).all()
    ref = torch.as_tensor(g["cine_diffusion_fixed"])
    assert abs(float(fast.abs().mean()) / float(ref.abs().mean()) - 1) < 0.2
    torch.set_grad_enabled(True)


def case_cine_diffusion_shift_graph(dev):
    """if_random_shift=True on the captured-graph path (per-step rolls read from a device table, ipdm_patch_fold_sched)
    against the per-step path that draws np.random.randint at every temporal step: same np.random seed, same Philox keys
    => the same chain; a different np.random seed => a different one (the rolls are really applied)."""
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
    n, T = 32, 8
    cfg = make_config("CINE127", 8, n, 10, 20.0, device=dev)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 5, cfg, dev)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1702, T, 1, n, n).to(dev)).reshape(4, 1, T, 1, n, n)
    cfg_T = make_config("CINE127", 128, T, 6, 0.2, device=dev)
    cfg_T.data.channels, cfg_T.data.channels_3d = 64, 1
    sig_T = get_sigmas(cfg_T)
    params = {"n_steps_each": 2, "step_lr": 1e-4}
    outs = {}
    for tag, graph, npseed in (("graph", True, 21), ("eager", False, 21), ("graph_other_rolls", True, 22)):
        net_T, _ = build_net(NCSN3DShallow, "NCSN3DShallow_ngf128", 13, cfg_T, dev)
        sampler = ALD.ALD2DTime(L2Penalty(A), net_T, sig_T, (1, T, 1, n, n), net, sig, params, cfg,
                                measurement=meas, linear_tfm=A, device=torch.device(dev))
        np.random.seed(npseed)
        outs[tag] = sampler(save_dir="/tmp", lr_scaled=1e4, mode_T="diffusion1d", lamda_T=0.5, if_random_shift=True, seed=3,
                            cuda_graph=graph)[0]
        torch.set_grad_enabled(True)
    assert torch.isfinite(outs["graph"].abs()).all()
    err = rel_l2(outs["graph"], outs["eager"])
    assert err < 1e-5, err
    assert rel_l2(outs["graph_other_rolls"], outs["graph"]) > 1e-4
    return err


# ------------------------------------------------------------------------------------------------ samplers
def case_sampler_uncond(dev):
    g = G("samplers")
    cfg = make_config("MNIST", 8, 28, 10, 20.0, device=dev)
    net, _ = build_net(NCSNv2, "NCSNv2_ngf8_28", 3, cfg, dev)
    sig = get_sigmas(cfg)
    params = {"n_steps_each": 2, "step_lr": 6.2e-6, "denoise": True, "final_only": True}
    draw = lambda shape: torch.randn(*shape)
    torch.manual_seed(101)
    res = ALD.ALDUnconditionalSampler((2, 1, 28, 28), net, sig, params, cfg, device=torch.device(dev))(noise_fn=draw)
    assert len(res) == 1 and res[0].device.type == "cpu"
    assert rel_l2(res[0], g["uncond_final"]) < TOL_X
    torch.manual_seed(101)
    x0 = torch.rand(2, 1, 28, 28).to(dev)
    res2 = anneal_Langevin_dynamics(x0, net, sig, 2, 6.2e-6, final_only=True, noise_fn=draw)
    assert rel_l2(res2[0], g["uncond_final"]) < TOL_X
    assert not torch.is_grad_enabled()      # Q12: the samplers leave grad mode off
    torch.set_grad_enabled(True)
    # sde 'ald' corrector
    sde = ns(N=10, T=1, marginal_prob=lambda x, t: (x, 0.01 * (20.0 / 0.01) ** t))
    score_fn = lambda x, t: net(x, torch.round((1 - t) * 9).long())
    torch.manual_seed(404)
    x = torch.rand(2, 1, 28, 28).to(dev)
    t = torch.tensor([0.6, 0.2]).to(dev)
    xo, xm = AnnealedLangevinDynamics(sde, score_fn, snr=0.176, n_steps=3).update_fn(x, t, noise_fn=draw)
    # three corrector steps whose update is score-dominated (step ~ 0.06, |x| ~ 1): the f16-operand
    # rounding of the score (~1e-3) shows up at ~1e-4 in x, so this case gets 5e-4
    assert rel_l2(xo.cpu(), g["sde_x"]) < 5e-4 and rel_l2(xm.cpu(), g["sde_mean"]) < 5e-4


def case_sampler_sense(dev):
    g = G("samplers")
    n = 32
    cfg = make_config("ACDC", 8, n, 10, 30.0, device=dev)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 4, cfg, dev)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    B = 2
    meas = A(phantom(1401, 1, 1, n, n).to(dev)).repeat(1, B, 1, 1, 1)
    params = {"n_steps_each": 2, "step_lr": 9e-7, "denoise": True, "final_only": True}
    draw = lambda shape: torch.randn(*shape)
    label = torch.zeros(B, 1, n, n, dtype=torch.long)
    for tag, lr_scaled in (("lr1e6", 1e6), ("lr1", 1.0)):
        sampler = ALD.ALDInvSegProximalRealImag(L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                                measurement=meas, linear_tfm=A, seg=None, device=torch.device(dev))
        torch.manual_seed(202)
        res = sampler(label=label, lamda=1.0, save_dir="/tmp", lr_scaled=lr_scaled, seg_mode="full", noise_fn=draw)
        assert res[0].dtype == torch.complex64 and tuple(res[0].shape) == (B, 1, n, n)
        assert rel_l2(res[0], g[f"sense_final_{tag}"]) < TOL_X, tag
    torch.set_grad_enabled(True)
    # the generic (non-fused) path through post_processing / self.proximal must agree with the fused kernel
    class Hooked(ALD.ALDInvSegProximalRealImag):
        def post_processing(self, xr, xi, **kw):
            return super().post_processing(xr, xi, **kw)
    sampler = Hooked(L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                     measurement=meas, linear_tfm=A, seg=None, device=torch.device(dev))
    torch.manual_seed(202)
    res = sampler(label=label, lamda=1.0, save_dir="/tmp", lr_scaled=1.0, seg_mode="full", noise_fn=draw)
    assert rel_l2(res[0], g["sense_final_lr1"]) < TOL_X
    torch.set_grad_enabled(True)


def case_sampler_cine(dev):
    g = G("samplers")
    n = 32
    cfg = make_config("CINE127", 8, n, 10, 20.0, device=dev)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 5, cfg, dev)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
    meas = A(phantom(1402, 24, 1, n, n).to(dev)).reshape(4, 1, 24, 1, n, n)
    net_T = ns(sigmas=None, config=ns(data=ns(channels=64)))
    sig_T = torch.tensor(np.exp(np.linspace(np.log(5.0), np.log(0.01), 6))).float().to(dev)
    params = {"n_steps_each": 1, "step_lr": 1e-4}
    draw = lambda shape: torch.randn(*shape)
    for mode_T in ("none", "tv"):
        sampler = ALD.ALD2DTime(L2Penalty(A), net_T, sig_T, (1, 24, 1, n, n), net, sig, params, cfg,
                                measurement=meas, linear_tfm=A, device=torch.device(dev))
        assert tuple(net_T.sigmas.shape) == tuple(sig.shape)      # Q14: temporal net's sigmas are overwritten
        torch.manual_seed(303)
        res = sampler(save_dir="/tmp", lr_scaled=1e4, mode_T=mode_T, lamda_T=0.05, if_random_shift=False, noise_fn=draw)
        assert tuple(res[0].shape) == (1, 24, 1, n, n)
        # script-default step_lr = 1e-4 makes the first levels score-dominated (step*|score| ~ |x|), so the
        # ~8e-4 f16-operand error of the score (the same for an f16-operand fp32-accumulate oracle,
        # oracle.scorenet.OPERAND_ROUND) reaches x at ~3e-4; the 1e-4 bar applies to cfg-2 steps (case above)
        assert rel_l2(res[0], g[f"cine_final_{mode_T}"]) < 1e-3, mode_T
    torch.set_grad_enabled(True)


def case_full_chain_metrics(dev, levels=40):
    """north_star: final NRMSE / SSIM of the posterior mean within 1e-3 of the reference.  A complete chain
    (sigma 30 -> 0.01 over `levels` levels x 3 steps + denoise) of 4 chains with identical injected noise on both
    sides; metrics of the mean magnitude image and the mean of per-chain metrics (helpers/visualizations.py:93,117-125)."""
    n, B = 32, 4
    cfg = make_config("ACDC", 8, n, levels, 30.0, device=dev)
    net, Pd = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 4, cfg, dev)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    truth = phantom(1401, 1, 1, n, n)
    meas = A(truth.to(dev)).repeat(1, B, 1, 1, 1)
    params = {"n_steps_each": 3, "step_lr": 9e-7, "denoise": True, "final_only": True}
    draw = lambda shape: torch.randn(*shape)
    sampler = ALD.ALDInvSegProximalRealImag(L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                            measurement=meas, linear_tfm=A, seg=None, device=torch.device(dev))
    torch.manual_seed(77)
    got = sampler(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", noise_fn=draw)[0]
    torch.set_grad_enabled(True)
    maps, mask = A.sens_maps, A.random_under_fourier.mask
    score = lambda x, y: SN.score_forward("NCSNv2Deepest", Pd, x, y)
    prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, mask, a, l)
    torch.manual_seed(77)
    with torch.no_grad():
        ref = OALD.ald_sense_real_imag(score, meas.cpu(), sig.cpu(), 3, 9e-7, 1e6, lambda s: M.sense_adjoint(s, maps), prox)
    t = truth.abs()[0, 0]
    rng = float(t.max() - t.min())
    def metrics(x):
        mags = x.abs()[:, 0]
        mean_img = mags.mean(0)
        per = [(OALD.nrmse(m, t), OALD.ssim(m, t, data_range=rng)) for m in mags]
        return (OALD.nrmse(mean_img, t), OALD.ssim(mean_img, t, data_range=rng),
                sum(p[0] for p in per) / len(per), sum(p[1] for p in per) / len(per))
    mg, mr = metrics(got), metrics(ref)
    for a, b in zip(mg, mr):
        assert abs(a - b) < 1e-3, (mg, mr)
    assert rel_l2(got, ref) < 1e-3
    return mg, mr


def case_map_baselines(dev):
    """MAP baselines on the same operators, against the reference's outputs (tests/golden/map.npz).
    Adam normalises the gradient, so a score error moves x by at most lr per step: tolerance 1e-3."""
    g = G("map")
    n = 32
    cfg = make_config("ACDC", 8, n, 10, 30.0, device=dev)
    cfg.MAP = ns(n_iters=6, lr=1e-2, complex_inner_n_steps=20)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 6, cfg, dev)
    A = SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1501, 1, 1, n, n).to(dev))
    x0 = A.conj_op(meas).clone()
    out = MAP.SENSEMAP(x0, meas, net, A, 0.5, cfg, None)()
    assert rel_l2(out.cpu(), g["map2d_final"]) < 1e-3
    cfg = make_config("CINE127", 8, n, 10, 20.0, device=dev)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 5, cfg, dev)
    A = SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
    meas = A(phantom(1402, 24, 1, n, n).to(dev)).reshape(4, 1, 24, 1, n, n)
    x0 = A.conj_op(meas.reshape(4, 24, 1, n, n)).reshape(1, 24, 1, n, n).clone()
    params = dict(lr=5e-3, opt_class=torch.optim.Adam, num_iters=3, num_plot_times=1, win_size=8, prior_weight=1.0,
                  spatial_step_weight=0.7, temporal_step_weight=0.05, save_dir="/tmp", opt_params={"betas": (0.5, 0.5)},
                  mode_T="tv", if_random_shift=False)
    rec = MAP.MAPOptimizer2DTime(x0, meas, net, None, A, None, params)()
    assert rel_l2(rec, g["map2dt_final"]) < 1e-3


def case_deeper_langevin_seg(dev):
    """NCSNv2Deeper forward, the 'langevin' corrector, and the segmentation-guidance hook (`adjust_grad`, run between
    score and update with torch autograd on the user's seg net) against the reference's outputs."""
    g = G("extra")
    cfg = make_config("ACDC", 8, 32, 12, 30.0)
    net, _ = build_net(NCSNv2Deeper, "NCSNv2Deeper_ngf8", 8, cfg, dev)
    out = net(rrand(1303, 2, 1, 32, 32).to(dev), torch.tensor([3, 10]).to(dev))
    assert rel_l2(out.cpu(), g["deeper_out"]) < TOL_SCORE
    cfg = make_config("MNIST", 8, 28, 10, 20.0, device=dev)
    net, _ = build_net(NCSNv2, "NCSNv2_ngf8_28", 3, cfg, dev)
    sde = ns(N=10, T=1)
    score_fn = lambda x, t: net(x, torch.round((1 - t) * 9).long())
    torch.manual_seed(405)
    x = torch.rand(2, 1, 28, 28).to(dev)
    xo, xm = LangevinCorrector(sde, score_fn, snr=0.16, n_steps=2).update_fn(x, torch.tensor([0.6, 0.2]).to(dev),
                                                                             noise_fn=lambda shape: torch.randn(*shape))
    assert rel_l2(xo.cpu(), g["lang_x"]) < 5e-4 and rel_l2(xm.cpu(), g["lang_mean"]) < 5e-4
    n, B = 32, 2
    cfg = make_config("ACDC", 8, n, 10, 30.0, device=dev)
    net, _ = build_net(NCSNv2Deepest, "NCSNv2Deepest_ngf8", 4, cfg, dev)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1401, 1, 1, n, n).to(dev)).repeat(1, B, 1, 1, 1)
    torch.manual_seed(9)
    seg = torch.nn.Conv2d(1, 3, 3, padding=1).to(dev)
    label = (rrand(1601, B, 1, n, n) * 3).long().clamp(max=2).to(dev)
    params = {"n_steps_each": 2, "step_lr": 9e-7, "denoise": True, "final_only": True}
    sampler = ALD.ALDInvSegProximalRealImag(L2Penalty(A), 0.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                            measurement=meas, linear_tfm=A, seg=seg, device=torch.device(dev))
    torch.manual_seed(203)
    res = sampler(label=label, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", noise_fn=lambda shape: torch.randn(*shape))
    torch.set_grad_enabled(True)
    assert rel_l2(res[0], g["seg_final"]) < TOL_X


def case_metrics(dev):
    """helpers.metrics / helpers.results against the oracle's restatement of the reference metrics (skimage
    conventions) on a batch of perturbed phantoms; incl. the reference's (B,C,H,W) / (1,C,H,W) call shape."""
    import tempfile
    from inverseproblemwithdiffusionmodel_b200.helpers import metrics as HM, results as HR
    B, n = 5, 72
    g = torch.Generator().manual_seed(77)
    orig = phantom(21, 1, 1, n, n)                                  # (1,1,H,W) complex
    recons = orig + 0.05 * torch.complex(torch.randn(B, 1, n, n, generator=g), torch.randn(B, 1, n, n, generator=g))
    mags, t = recons.abs(), orig.abs()
    rng = float(t.max() - t.min())
    got = HM.compute_metrics(["NRMSE", "SSIM", "L2", "L1"], mags.to(dev), t.to(dev), data_range=rng)
    for i in range(B):
        assert abs(got["NRMSE"][i] - OALD.nrmse(mags[i, 0], t[0, 0])) < 1e-6
        assert abs(got["SSIM"][i] - OALD.ssim(mags[i, 0], t[0, 0], data_range=rng)) < 1e-6
        assert abs(got["L2"][i] - float(((mags[i] - t[0]) ** 2).mean())) < 1e-8
        assert abs(got["L1"][i] - float((mags[i] - t[0]).abs().mean())) < 1e-7
    red = HM.compute_metrics(["NRMSE"], mags.to(dev), t.to(dev), reduce="mean")
    assert abs(red["NRMSE"] - got["NRMSE"].mean()) < 1e-12
    # multi-channel SSIM = mean of the per-channel values; per-image references
    two = torch.cat([mags, mags.flip(-1)], 1)
    ref2 = torch.cat([t, t.flip(-1)], 1).repeat(B, 1, 1, 1) + 0.01
    s2 = HM.SSIM_wrapper(two.to(dev), ref2.to(dev), data_range=rng)
    for i in range(B):
        want = 0.5 * (OALD.ssim(two[i, 0], ref2[i, 0], data_range=rng) + OALD.ssim(two[i, 1], ref2[i, 1], data_range=rng))
        assert abs(s2[i] - want) < 1e-6
    mm, pm, ms, ps = HM.compute_mean_and_std(recons.to(dev))
    st = OALD.posterior_stats(recons)
    assert rel_l2(mm.cpu(), st["mag_mean"]) < 1e-6 and rel_l2(ms.cpu(), st["mag_std"]) < 1e-4
    assert rel_l2(pm.cpu(), st["phase_mean"]) < 1e-5 and rel_l2(ps.cpu(), st["phase_std"]) < 1e-4
    snr = HM.compute_snr(recons.to(dev))
    flat = recons.abs().reshape(B, -1)
    assert np.allclose(snr, (20 * torch.log10(flat.max(1).values / flat.std(1, unbiased=False))).numpy(), rtol=1e-5)
    with tempfile.TemporaryDirectory() as d:
        HR.save_reconstruction(d, orig, torch.zeros(4, 1, 1, n, n, dtype=torch.complex64), recons.to(dev), zero_filled=orig[0],
                               args_dict={"R": 40})
        back = HR.load_reconstruction(d)
        assert sorted(os.listdir(d)) == ["ZF.pt", "args_dict.pkl", "measurement.pt", "original.pt", "reconstructions.pt"]
        assert torch.equal(back["reconstructions"], recons) and back["reconstructions"].device.type == "cpu"
        assert back["args_dict"] == {"R": 40}
