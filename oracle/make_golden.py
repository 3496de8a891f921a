"""TEST INFRASTRUCTURE -- generates `tests/golden/*.npz` by running the REAL reference.

Run in the authoring container only (needs /root/reference):  `python -m oracle.make_golden`
The fixtures are outputs of the unmodified reference classes (imported through
`oracle/ref_shim.py`) on small seeded inputs; `tests/test_oracle_golden.py` pins the oracle
restatement to them, `tests/test_gpu_*.py` pin the CUDA path to the oracle and to them.
"""
import argparse
import io
import json
import os
import contextlib
import types

import numpy as np
import torch

from oracle import ref_shim
from oracle.scorenet import synth_state_dict
from oracle.fixture_inputs import crandn, rrand, rrandn, phantom

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def small_cfg(name, ngf, image_size, num_classes, sigma_begin, sigma_end=0.01):
    cfg = ref_shim.load_config(name)
    cfg.model.ngf = ngf
    cfg.data.image_size = image_size
    cfg.model.num_classes = num_classes
    cfg.model.sigma_begin = sigma_begin
    cfg.model.sigma_end = sigma_end
    if hasattr(cfg, "recons"):
        cfg.recons.num_classes = num_classes
        cfg.recons.sigma_begin = sigma_begin
        cfg.recons.sigma_end = sigma_end
    return cfg


def build_ref_net(cls_name, cfg, seed):
    from InverseProblemWithDiffusionModel.ncsn.models import ncsnv2

    net = getattr(ncsnv2, cls_name)(cfg).eval()
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    P = synth_state_dict(spec, seed, net.sigmas)
    net.load_state_dict(P)
    return net, spec, P


def gen_fft_mask_coils():
    from InverseProblemWithDiffusionModel.ncsn.linear_transforms import i2k_complex, k2i_complex, generate_mask
    from InverseProblemWithDiffusionModel.ncsn.linear_transforms.undersampling_fourier import SENSE, RandomUndersamplingFourier

    out = {}
    for n in (16, 32):
        x = crandn(1100 + n, 2, 1, n, n)
        out[f"fft_i2k_{n}"] = _np(i2k_complex(x))
        out[f"fft_k2i_{n}"] = _np(k2i_complex(x))
    # rectangular + real input
    x = rrandn(1199, 1, 1, 8, 32)
    out["fft_i2k_rect"] = _np(i2k_complex(x))
    out["fft_k2i_rect"] = _np(k2i_complex(x))

    out["mask_R8_T24_N128_seed3"] = _np(generate_mask(24, 128, sw=0.196, sm=0.5, sa=0.02, seed=3))
    out["mask_T1_N64_seed5"] = _np(generate_mask(1, 64, seed=5))
    for W in (32, 128, 256):
        out[f"mask_live_W{W}_seed0"] = _np(RandomUndersamplingFourier(40, 1 / 64, (1, W, W), seed=0).mask)
    A = SENSE("exp", 4, 40, 1 / 64, (1, 32, 32), 0)
    out["coil_maps_32_seed0"] = _np(A.sens_maps)
    for n in (128, 256):
        A2 = SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
        m = A2.sens_maps
        out[f"coil_maps_{n}_seed0_stats"] = np.array([float(m.sum()), float(m.min()), float(m.max()),
                                                       float((m ** 2).sum()), float(m[1, 17, 101]), float(m[3, n - 1, 5])])
    np.savez_compressed(os.path.join(OUT, "linear_ops.npz"), **out)


def gen_sense_prox():
    from InverseProblemWithDiffusionModel.ncsn.linear_transforms.undersampling_fourier import SENSE, RandomUndersamplingFourier
    from InverseProblemWithDiffusionModel.ncsn.models.proximal_op import L2Penalty, SingleCoil
    from oracle.mri_ops import keep_center_mask

    out = {}
    n = 32
    A = SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
    # (i) the live (24,1,1,W) mask, batch 24 (16x16 frames, 3 coils, to keep the fixture small)
    A2c = SENSE("exp", 2, 40, 1 / 64, (1, n, n), 7)
    x24 = crandn(1201, 24, 1, n, n)
    S24 = A2c(x24)
    out["live_S_frames_0_5_23"] = _np(S24[:, [0, 5, 23]])
    out["live_adj"] = _np(A2c.conj_op(S24))
    out["live_ssos"] = _np(A2c.SSOS(S24))
    # conj_op on an UNMASKED input (quirk Q3: no mask applied)
    y_dense = crandn(1202, 4, 2, 1, n, n)
    out["dense_adj"] = _np(A.conj_op(y_dense))
    # (ii) overwritten keep-centre (1,1,W) mask, arbitrary batch
    kc = keep_center_mask(n, 4, 1 / 8, seed=0)
    A.random_under_fourier.mask = kc
    x3 = crandn(1203, 3, 1, n, n)
    S3 = A(x3)
    out["kc_mask"], out["kc_S"] = _np(kc), _np(S3)
    out["kc_adj"] = _np(A.conj_op(S3))
    out["kc_loglh"] = _np(A.log_lh_grad(x3, S3 * 0.5, 0.7))
    # L2Penalty through the reference's autograd + SGD implementation
    z = crandn(1204, 3, 1, n, n)
    prox = L2Penalty(A)
    for tag, alpha in (("a1", 1.0), ("a1e3", 1e3)):
        with _quiet():
            out[f"l2_{tag}"] = _np(prox(z, S3, alpha, 1.0))
    torch.set_grad_enabled(True)
    # single-coil operator + exact prox + projection
    F1 = RandomUndersamplingFourier(4, 1 / 8, (1, n, n), seed=0)
    F1.mask = kc
    S1 = F1(x3)
    out["sc_S"] = _np(S1)
    with _quiet():
        out["sc_prox"] = _np(SingleCoil(F1)(z, S1, 0.8, 1.0))
        out["sc_l2"] = _np(L2Penalty(F1)(z, S1, 2.0, 1.0))
    torch.set_grad_enabled(True)
    out["sc_proj"] = _np(F1.projection(z, S1, 0.3))
    np.savez_compressed(os.path.join(OUT, "sense_prox.npz"), **out)


def gen_scorenet():
    out = {}
    specs = {}
    with torch.no_grad():
        cfg = small_cfg("acdc", 8, 32, 12, 30.0)
        net, spec, _ = build_ref_net("NCSNv2Deepest", cfg, seed=1)
        specs["NCSNv2Deepest_ngf8"] = [[k, list(s)] for k, s in spec]
        x = rrand(1301, 2, 1, 32, 32) * 3 - 1
        y = torch.tensor([0, 7])
        out["deepest_out"] = _np(net(x, y))
        cfg = small_cfg("mnist", 8, 28, 12, 30.0)
        net, spec, _ = build_ref_net("NCSNv2", cfg, seed=2)
        specs["NCSNv2_ngf8_28"] = [[k, list(s)] for k, s in spec]
        x = rrand(1302, 2, 1, 28, 28)
        y = torch.tensor([11, 3])
        out["v2_out"] = _np(net(x, y))
        # full-width key/shape census (layout contract for load_state_dict)
        cfg = ref_shim.load_config("acdc")
        from InverseProblemWithDiffusionModel.ncsn.models import ncsnv2
        full = ncsnv2.NCSNv2Deepest(cfg)
        specs["NCSNv2Deepest_acdc"] = [[k, list(v.shape)] for k, v in full.state_dict().items()]
        cfg = ref_shim.load_config("mnist")
        cfg.data.image_size = 28
        full = ncsnv2.NCSNv2(cfg)
        specs["NCSNv2_mnist28"] = [[k, list(v.shape)] for k, v in full.state_dict().items()]
    np.savez_compressed(os.path.join(OUT, "scorenet.npz"), **out)
    with open(os.path.join(OUT, "state_dict_specs.json"), "w") as f:
        json.dump(specs, f)


def gen_samplers():
    from InverseProblemWithDiffusionModel.ncsn.models import ALD_optimizers as ALD
    from InverseProblemWithDiffusionModel.ncsn.models import anneal_Langevin_dynamics, get_sigmas
    from InverseProblemWithDiffusionModel.ncsn.models.proximal_op import L2Penalty
    from InverseProblemWithDiffusionModel.ncsn.linear_transforms.undersampling_fourier import SENSE
    from oracle.mri_ops import keep_center_mask

    out = {}
    # ---- cfg 1 shaped: unconditional ALD with NCSNv2 on 28x28 ------------------------------------
    cfg = small_cfg("mnist", 8, 28, 10, 20.0)
    net, _, _ = build_ref_net("NCSNv2", cfg, seed=3)
    sig = get_sigmas(cfg)
    params = {"n_steps_each": 2, "step_lr": 6.2e-6, "denoise": True, "final_only": True}
    torch.manual_seed(101)
    with _quiet():
        res = ALD.ALDUnconditionalSampler((2, 1, 28, 28), net, sig, params, cfg, device=torch.device("cpu"))()
    out["uncond_final"] = _np(res[0])
    torch.manual_seed(101)
    x0 = torch.rand(2, 1, 28, 28)
    with _quiet():
        res2 = anneal_Langevin_dynamics(x0, net, sig, 2, 6.2e-6, final_only=True)
    assert torch.equal(res[0], res2[0]), "class sampler and functional ALD must agree bit for bit"
    torch.set_grad_enabled(True)

    # ---- cfg 2 shaped: SENSE real/imag prox ALD with NCSNv2Deepest on 32x32 ----------------------
    n = 32
    cfg = small_cfg("acdc", 8, n, 10, 30.0)
    net, _, _ = build_ref_net("NCSNv2Deepest", cfg, seed=4)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    img = phantom(1401, 1, 1, n, n)
    B = 2
    meas = A(img).repeat(1, B, 1, 1, 1)
    seg = torch.nn.Conv2d(1, 2, 3, padding=1)
    label = torch.zeros(B, 1, n, n, dtype=torch.long)
    params = {"n_steps_each": 2, "step_lr": 9e-7, "denoise": True, "final_only": True}
    for tag, lr_scaled in (("lr1e6", 1e6), ("lr1", 1.0)):
        sampler = ALD.ALDInvSegProximalRealImag(L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                                measurement=meas, linear_tfm=A, seg=seg, device=torch.device("cpu"))
        torch.manual_seed(202)
        with _quiet():
            res = sampler(label=label, lamda=1.0, save_dir="/tmp/ipdm_golden", lr_scaled=lr_scaled, seg_mode="full")
        out[f"sense_final_{tag}"] = _np(res[0])
        torch.set_grad_enabled(True)

    # ---- cfg 4 shaped: 2D+time with the live (24,1,1,W) mask, 32x32 frames -----------------------
    n = 32
    cfg = small_cfg("cine127", 8, n, 10, 20.0)
    net, _, _ = build_ref_net("NCSNv2Deepest", cfg, seed=5)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
    vol = phantom(1402, 24, 1, n, n)
    meas = A(vol).reshape(4, 1, 24, 1, n, n)
    net_T = types.SimpleNamespace(sigmas=None, config=types.SimpleNamespace(data=types.SimpleNamespace(channels=64)))
    sig_T = torch.tensor(np.exp(np.linspace(np.log(5.0), np.log(0.01), 6))).float()
    params = {"n_steps_each": 1, "step_lr": 1e-4}
    for mode_T in ("none", "tv"):
        sampler = ALD.ALD2DTime(L2Penalty(A), net_T, sig_T, (1, 24, 1, n, n), net, sig, params, cfg,
                                measurement=meas, linear_tfm=A, device=torch.device("cpu"))
        torch.manual_seed(303)
        with _quiet():
            res = sampler(save_dir="/tmp/ipdm_golden", lr_scaled=1e4, mode_T=mode_T, lamda_T=0.05, if_random_shift=False)
        out[f"cine_final_{mode_T}"] = _np(res[0])
        torch.set_grad_enabled(True)

    # ---- sde 'ald' corrector --------------------------------------------------------------------
    from InverseProblemWithDiffusionModel.sde import sampling as sde_sampling, sde_lib
    sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=20.0, N=10)
    cfg = small_cfg("mnist", 8, 28, 10, 20.0)
    net, _, _ = build_ref_net("NCSNv2", cfg, seed=3)

    def score_fn(x, t):
        labels = torch.round((sde.T - t) * (sde.N - 1)).long()
        return net(x, labels)

    corr = sde_sampling.AnnealedLangevinDynamics(sde, score_fn, snr=0.176, n_steps=3)
    torch.manual_seed(404)
    x = torch.rand(2, 1, 28, 28)
    t = torch.tensor([0.6, 0.2])
    with torch.no_grad():
        xo, xm = corr.update_fn(x, t)
    out["sde_x"], out["sde_mean"] = _np(xo), _np(xm)
    np.savez_compressed(os.path.join(OUT, "samplers.npz"), **out)


def gen_extra():
    """NCSNv2Deeper forward, the 'langevin' corrector, segmentation-guided ALD."""
    from InverseProblemWithDiffusionModel.ncsn.models import ALD_optimizers as ALD
    from InverseProblemWithDiffusionModel.ncsn.models import get_sigmas
    from InverseProblemWithDiffusionModel.ncsn.models.proximal_op import L2Penalty
    from InverseProblemWithDiffusionModel.ncsn.linear_transforms.undersampling_fourier import SENSE
    from InverseProblemWithDiffusionModel.sde import sampling as sde_sampling, sde_lib
    from oracle.mri_ops import keep_center_mask
    out = {}
    specs_path = os.path.join(OUT, "state_dict_specs.json")
    specs = json.load(open(specs_path))
    with torch.no_grad():
        cfg = small_cfg("acdc", 8, 32, 12, 30.0)
        net, spec, _ = build_ref_net("NCSNv2Deeper", cfg, seed=8)
        specs["NCSNv2Deeper_ngf8"] = [[k, list(s)] for k, s in spec]
        out["deeper_out"] = _np(net(rrand(1303, 2, 1, 32, 32), torch.tensor([3, 10])))
    json.dump(specs, open(specs_path, "w"))
    # 'langevin' corrector
    sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=20.0, N=10)
    cfg = small_cfg("mnist", 8, 28, 10, 20.0)
    net, _, _ = build_ref_net("NCSNv2", cfg, seed=3)
    score_fn = lambda x, t: net(x, torch.round((sde.T - t) * (sde.N - 1)).long())
    corr = sde_sampling.LangevinCorrector(sde, score_fn, snr=0.16, n_steps=2)
    torch.manual_seed(405)
    x = torch.rand(2, 1, 28, 28)
    with torch.no_grad():
        xo, xm = corr.update_fn(x, torch.tensor([0.6, 0.2]))
    out["lang_x"], out["lang_mean"] = _np(xo), _np(xm)
    # segmentation-guided chain: weights ramp from level 0 (seg_start_time = 0), tiny seeded conv net as `seg`
    n, B = 32, 2
    cfg = small_cfg("acdc", 8, n, 10, 30.0)
    net, _, _ = build_ref_net("NCSNv2Deepest", cfg, seed=4)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1401, 1, 1, n, n)).repeat(1, B, 1, 1, 1)
    torch.manual_seed(9)
    seg = torch.nn.Conv2d(1, 3, 3, padding=1)
    label = (rrand(1601, B, 1, n, n) * 3).long().clamp(max=2)
    params = {"n_steps_each": 2, "step_lr": 9e-7, "denoise": True, "final_only": True}
    sampler = ALD.ALDInvSegProximalRealImag(L2Penalty(A), 0.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                            measurement=meas, linear_tfm=A, seg=seg, device=torch.device("cpu"))
    torch.manual_seed(203)
    with _quiet():
        res = sampler(label=label, lamda=1.0, save_dir="/tmp/ipdm_golden", lr_scaled=1e6, seg_mode="full")
    torch.set_grad_enabled(True)
    out["seg_final"] = _np(res[0])
    np.savez_compressed(os.path.join(OUT, "extra.npz"), **out)


def gen_map():
    """MAP baselines: reference MAPOptimizer / MAPOptimizer2DTime with a no-op logger."""
    import importlib
    MAP = importlib.import_module(ref_shim.PKG + ".ncsn.models.MAP_optimizers")
    from InverseProblemWithDiffusionModel.ncsn.linear_transforms.undersampling_fourier import SENSE
    from oracle.mri_ops import keep_center_mask
    logger = types.SimpleNamespace(add_scalar=lambda *a, **k: None, add_image=lambda *a, **k: None)
    out = {}
    n = 32
    cfg = small_cfg("acdc", 8, n, 10, 30.0)
    cfg.MAP.n_iters = 50          # plot_interval = n_iters // 50 must be >= 1
    cfg.MAP.lr = 1e-2
    net, _, _ = build_ref_net("NCSNv2Deepest", cfg, seed=6)
    A = SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1501, 1, 1, n, n))
    x0 = A.conj_op(meas).clone()
    with _quiet():
        opt = MAP.SENSEMAP(x0, meas, net, A, 0.5, cfg, logger, device=torch.device("cpu"))
        # 6 iterations are enough to pin the arithmetic; drive _step directly (the __call__ loop only adds logging)
        with torch.no_grad():
            x = x0
            for it in range(6):
                x = opt._step(x, it)
    out["map2d_final"] = _np(x)
    # 2D+time with the TV temporal term
    cfg = small_cfg("cine127", 8, n, 10, 20.0)
    net, _, _ = build_ref_net("NCSNv2Deepest", cfg, seed=5)
    A = SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
    vol = phantom(1402, 24, 1, n, n)
    meas = A(vol).reshape(4, 1, 24, 1, n, n)
    x0 = A.conj_op(meas.reshape(4, 24, 1, n, n)).reshape(1, 24, 1, n, n).clone()
    net_T = types.SimpleNamespace(config=types.SimpleNamespace(data=types.SimpleNamespace(channels=64)))
    params = dict(lr=5e-3, opt_class=torch.optim.Adam, num_iters=3, num_plot_times=1, win_size=8, prior_weight=1.0,
                  spatial_step_weight=0.7, temporal_step_weight=0.05, save_dir="/tmp/ipdm_golden", opt_params={"betas": (0.5, 0.5)},
                  mode_T="tv", if_random_shift=False, device=torch.device("cpu"))
    MAP.save_vol_as_gif = lambda *a, **k: None
    MAP.vis_images = lambda *a, **k: None
    MAP.vis_multi_channel_signal = lambda *a, **k: None
    MAP.normalize_phase = lambda x: x
    with _quiet():
        rec = MAP.MAPOptimizer2DTime(x0, meas, net, net_T, A, logger, params)()
    out["map2dt_final"] = _np(rec)
    np.savez_compressed(os.path.join(OUT, "map.npz"), **out)


def gen_ncsn3d():
    """NCSN3DShallow, the learned temporal prior: full width (ngf 128, what the tensor-core path
    needs), two 8x8x24 patches, 5-D and flattened 3-D input conventions."""
    from InverseProblemWithDiffusionModel.ncsn.models import ncsn3d
    out = {}
    specs_path = os.path.join(OUT, "state_dict_specs.json")
    specs = json.load(open(specs_path))
    with torch.no_grad():
        cfg = small_cfg("cine127_1d", 128, 24, 12, 40.0)
        net = ncsn3d.NCSN3DShallow(cfg).eval()
        spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
        P = synth_state_dict(spec, 12, net.sigmas)
        net.load_state_dict(P)
        specs["NCSN3DShallow_ngf128"] = [[k, list(s)] for k, s in spec]
        x = rrand(1701, 2, 1, 8, 8, 24)
        y = torch.tensor([2, 9])
        out["shallow_out"] = _np(net(x, y))
        out["shallow_out_flat"] = _np(net(x.reshape(2, 64, 24), y))
    json.dump(specs, open(specs_path, "w"))
    # ---- 2D+time chain with the learned temporal prior: 32x32 frames, T = 8, 4 coils, (1,1,W) mask; the temporal
    # schedule starts at 0.2 so that the four last spatial levels run the temporal step and the first six skip it
    from InverseProblemWithDiffusionModel.ncsn.models import ALD_optimizers as ALD
    from InverseProblemWithDiffusionModel.ncsn.models import get_sigmas
    from InverseProblemWithDiffusionModel.ncsn.models.proximal_op import L2Penalty
    from InverseProblemWithDiffusionModel.ncsn.linear_transforms.undersampling_fourier import SENSE
    from oracle.mri_ops import keep_center_mask
    n, T = 32, 8
    cfg = small_cfg("cine127", 8, n, 10, 20.0)
    net, _, _ = build_ref_net("NCSNv2Deepest", cfg, seed=5)
    sig = get_sigmas(cfg, mode="recons")
    A = SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1702, T, 1, n, n)).reshape(4, 1, T, 1, n, n)
    cfg_T = small_cfg("cine127_1d", 128, T, 6, 0.2)
    sig_T = get_sigmas(cfg_T)
    params = {"n_steps_each": 1, "step_lr": 1e-4}
    for tag, shift in (("fixed", False), ("shift", True)):
        net_T = ncsn3d.NCSN3DShallow(cfg_T).eval()
        net_T.load_state_dict(synth_state_dict([(k, tuple(v.shape)) for k, v in net_T.state_dict().items()], 13, net_T.sigmas))
        sampler = ALD.ALD2DTime(L2Penalty(A), net_T, sig_T, (1, T, 1, n, n), net, sig, params, cfg,
                                measurement=meas, linear_tfm=A, device=torch.device("cpu"))
        torch.manual_seed(304)
        np.random.seed(11)
        with _quiet():
            res = sampler(save_dir="/tmp/ipdm_golden", lr_scaled=1e4, mode_T="diffusion1d", lamda_T=0.5, if_random_shift=shift)
        out[f"cine_diffusion_{tag}"] = _np(res[0])
        torch.set_grad_enabled(True)
    # ---- MAP baseline with the learned temporal prior (MAPOptimizer2DTime, mode_T = "diffusion1d"), same data ----
    import importlib
    MAP = importlib.import_module(ref_shim.PKG + ".ncsn.models.MAP_optimizers")
    logger = types.SimpleNamespace(add_scalar=lambda *a, **k: None, add_image=lambda *a, **k: None)
    MAP.save_vol_as_gif = lambda *a, **k: None
    MAP.vis_images = lambda *a, **k: None
    MAP.vis_multi_channel_signal = lambda *a, **k: None
    MAP.normalize_phase = lambda x: x
    x0 = A.conj_op(meas.reshape(4, T, 1, n, n)).reshape(1, T, 1, n, n).clone()
    net_T = ncsn3d.NCSN3DShallow(cfg_T).eval()
    net_T.load_state_dict(synth_state_dict([(k, tuple(v.shape)) for k, v in net_T.state_dict().items()], 13, net_T.sigmas))
    params = dict(lr=5e-3, opt_class=torch.optim.Adam, num_iters=2, num_plot_times=1, win_size=8, prior_weight=1.0,
                  spatial_step_weight=0.7, temporal_step_weight=0.3, save_dir="/tmp/ipdm_golden", opt_params={"betas": (0.5, 0.5)},
                  mode_T="diffusion1d", if_random_shift=False, device=torch.device("cpu"))
    with _quiet():
        rec = MAP.MAPOptimizer2DTime(x0, meas, net, net_T, A, logger, params)()
    out["map2dt_diffusion"] = _np(rec)
    torch.set_grad_enabled(True)
    np.savez_compressed(os.path.join(OUT, "ncsn3d.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    ref_shim.install()
    os.makedirs(OUT, exist_ok=True)
    os.makedirs("/tmp/ipdm_golden", exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    todo = {"linear": gen_fft_mask_coils, "sense": gen_sense_prox, "scorenet": gen_scorenet, "samplers": gen_samplers,
            "map": gen_map, "extra": gen_extra, "ncsn3d": gen_ncsn3d}
    for name, fn in todo.items():
        if args.only and args.only != name:
            continue
        fn()
        print("wrote", name)


if __name__ == "__main__":
    main()
