"""Pins the oracle restatement (oracle/) to outputs of the REAL reference (tests/golden/, made by
oracle/make_golden.py in the authoring container).  CPU only.  Tolerances: integer/bool work is
bit-exact; fp32 work is <= 2e-6 relative L2 (same torch CPU kernels, different op order at most)."""
import json
import os

import numpy as np
import torch

from conftest import rel_l2, GOLDEN
from oracle import mri_ops as M
from oracle import ald as ALD
from oracle import scorenet as SN
from oracle.fixture_inputs import crandn, rrand, rrandn, phantom

TOL = 2e-6


def test_centered_fft(golden):
    G = golden("linear_ops")
    for n in (16, 32):
        x = crandn(1100 + n, 2, 1, n, n)
        assert rel_l2(M.i2k(x), G[f"fft_i2k_{n}"]) < TOL
        assert rel_l2(M.k2i(x), G[f"fft_k2i_{n}"]) < TOL
    x = rrandn(1199, 1, 1, 8, 32)
    assert rel_l2(M.i2k(x), G["fft_i2k_rect"]) < TOL
    assert rel_l2(M.k2i(x), G["fft_k2i_rect"]) < TOL


def test_checkerboard_identity():
    """i2k(x) == P * fft2_ortho(P * x) * (-1)^(H/2+W/2), P = (-1)^(i+j), even sizes."""
    for (h, w) in ((16, 16), (8, 32), (32, 64)):
        x = crandn(7, 2, h, w)
        ii, jj = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        P = (1 - 2 * ((ii + jj) % 2)).to(torch.complex64)
        sign = (-1) ** (h // 2 + w // 2)
        assert rel_l2(sign * P * torch.fft.fft2(P * x, norm="ortho"), M.i2k(x)) < TOL
        assert rel_l2(sign * P * torch.fft.ifft2(P * x, norm="ortho"), M.k2i(x)) < TOL


def test_masks_bit_exact(golden):
    G = golden("linear_ops")
    assert np.array_equal(M.variable_density_masks(24, 128, sw=0.196, sm=0.5, sa=0.02, seed=3).numpy(), G["mask_R8_T24_N128_seed3"])
    assert np.array_equal(M.variable_density_masks(1, 64, seed=5).numpy(), G["mask_T1_N64_seed5"])
    for W in (32, 128, 256):
        m = M.live_sense_mask(W, 0)
        assert m.shape == (24, 1, 1, W) and m.dtype == torch.bool
        assert np.array_equal(m.numpy(), G[f"mask_live_W{W}_seed0"])


def test_coil_maps(golden):
    G = golden("linear_ops")
    m = M.exp_coil_maps(4, 32, 32, 0)
    assert m.dtype == torch.float64
    assert np.allclose(m.numpy(), G["coil_maps_32_seed0"], rtol=0, atol=1e-14)
    assert torch.allclose((m ** 2).sum(0), torch.ones(32, 32, dtype=torch.float64))
    for n in (128, 256):
        m = M.exp_coil_maps(4, n, n, 0)
        stats = np.array([float(m.sum()), float(m.min()), float(m.max()), float((m ** 2).sum()),
                          float(m[1, 17, 101]), float(m[3, n - 1, 5])])
        assert np.allclose(stats, G[f"coil_maps_{n}_seed0_stats"], rtol=1e-12)


def test_sense_live_mask(golden):
    G = golden("sense_prox")
    n = 32
    maps = M.exp_coil_maps(2, n, n, 7)
    mask = M.live_sense_mask(n, 7)
    x = crandn(1201, 24, 1, n, n)
    S = M.sense_forward(x, maps, mask)
    assert S.shape == (2, 24, 1, n, n) and S.dtype == torch.complex64
    assert rel_l2(S[:, [0, 5, 23]], G["live_S_frames_0_5_23"]) < TOL
    assert rel_l2(M.sense_adjoint(S, maps), G["live_adj"]) < TOL
    assert rel_l2(M.sense_ssos(S), G["live_ssos"]) < TOL


def test_sense_keep_center_and_prox(golden):
    G = golden("sense_prox")
    n = 32
    maps = M.exp_coil_maps(4, n, n, 0)
    assert rel_l2(M.sense_adjoint(crandn(1202, 4, 2, 1, n, n), maps), G["dense_adj"]) < TOL
    kc = M.keep_center_mask(n, 4, 1 / 8, seed=0)
    assert np.array_equal(kc.numpy(), G["kc_mask"])
    x3 = crandn(1203, 3, 1, n, n)
    S3 = M.sense_forward(x3, maps, kc)
    assert rel_l2(S3, G["kc_S"]) < TOL
    assert rel_l2(M.sense_adjoint(S3, maps), G["kc_adj"]) < TOL
    fwd = lambda v: M.sense_forward(v, maps, kc)
    adj = lambda s: M.sense_adjoint(s, maps)
    assert rel_l2(M.log_lh_grad(fwd, adj, x3, S3 * 0.5, 0.7), G["kc_loglh"]) < TOL
    z = crandn(1204, 3, 1, n, n)
    for tag, alpha in (("a1", 1.0), ("a1e3", 1e3)):
        ref = torch.as_tensor(G[f"l2_{tag}"])
        assert rel_l2(M.l2_prox_sgd(fwd, z, S3, alpha, 1.0), ref) < TOL
        # the closed form moves z by the same amount as the autograd/SGD step
        cf = M.l2_prox_sense_closed_form(z, S3, maps, kc, alpha, 1.0)
        assert rel_l2(cf - z, ref - z) < 2e-5
        assert rel_l2(cf, ref) < TOL
    # single coil
    S1 = M.undersampled_fourier(x3, kc)
    assert rel_l2(S1, G["sc_S"]) < TOL
    assert rel_l2(M.single_coil_prox(z, S1, kc, 0.8, 1.0), G["sc_prox"]) < TOL
    f1 = lambda v: M.undersampled_fourier(v, kc)
    assert rel_l2(M.l2_prox_sgd(f1, z, S1, 2.0, 1.0), G["sc_l2"]) < TOL
    assert rel_l2(M.fourier_projection(z, S1, kc, 0.3), G["sc_proj"]) < TOL
    # analytic cross-check: the exact single-coil prox satisfies its normal equations
    xs = M.single_coil_prox(z, S1, kc, 0.8, 1.0)
    assert float(M.prox_residual(f1, M.k2i, xs, z, S1, 0.8, 1.0)) < 1e-8


def test_adjoint_dot_product():
    """<A x, y> == <x, A^H y> for masked y (conj_op is the adjoint only on masked input, Q3)."""
    n = 32
    maps = M.exp_coil_maps(4, n, n, 0)
    kc = M.keep_center_mask(n, 4, 1 / 8, seed=0)
    x = crandn(1, 2, 1, n, n)
    y = kc * crandn(2, 4, 2, 1, n, n)
    lhs = torch.vdot(y.reshape(-1), M.sense_forward(x, maps, kc).reshape(-1))
    rhs = torch.vdot(M.sense_adjoint(y, maps).reshape(-1), x.reshape(-1))
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


def _specs():
    with open(os.path.join(GOLDEN, "state_dict_specs.json")) as f:
        return json.load(f)


def _net(spec_name, seed, sigmas):
    spec = [(k, tuple(s)) for k, s in _specs()[spec_name]]
    return SN.synth_state_dict(spec, seed, sigmas)


def test_scorenet_forward(golden):
    G = golden("scorenet")
    sig = ALD.geometric_sigmas(30.0, 0.01, 12)
    with torch.no_grad():
        P = _net("NCSNv2Deepest_ngf8", 1, sig)
        out = SN.score_forward("NCSNv2Deepest", P, rrand(1301, 2, 1, 32, 32) * 3 - 1, torch.tensor([0, 7]))
        assert rel_l2(out, G["deepest_out"]) < 1e-5
        P = _net("NCSNv2_ngf8_28", 2, sig)
        out = SN.score_forward("NCSNv2", P, rrand(1302, 2, 1, 28, 28), torch.tensor([11, 3]))
        assert rel_l2(out, G["v2_out"]) < 1e-5


def test_ncsn3d_shallow_forward(golden):
    """oracle restatement of NCSN3DShallow (temporal prior) pinned to the reference's output at full width (ngf 128)."""
    G = golden("ncsn3d")
    sig = ALD.geometric_sigmas(40.0, 0.01, 12)
    with torch.no_grad():
        P = _net("NCSN3DShallow_ngf128", 12, sig)
        x = rrand(1701, 2, 1, 8, 8, 24)
        out = SN.score_forward_3d_shallow(P, x, torch.tensor([2, 9]))
    assert rel_l2(out, G["shallow_out"]) < 1e-5
    assert rel_l2(G["shallow_out_flat"].reshape(2, 1, 8, 8, 24), G["shallow_out"]) < 1e-6


def test_cine_chain_with_temporal_prior(golden):
    """oracle ALD2DTime mode_T='diffusion1d' (fold, remapped sigma_T, temporal Langevin step, unfold, np.random rolls)
    pinned to the reference's chains."""
    G = golden("ncsn3d")
    n, T = 32, 8
    sig = ALD.geometric_sigmas(20.0, 0.01, 10)
    sig_T = ALD.remap_sigmas_T(sig, ALD.geometric_sigmas(0.2, 0.01, 6))
    maps, mask = M.exp_coil_maps(4, n, n, 0), M.keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = M.sense_forward(phantom(1702, T, 1, n, n), maps, mask).reshape(4, 1, T, 1, n, n)
    P2 = _net("NCSNv2Deepest_ngf8", 5, sig)
    P3 = _net("NCSN3DShallow_ngf128", 13, ALD.geometric_sigmas(0.2, 0.01, 6))
    P3["sigmas"] = sig_T                                            # Q14
    score = lambda x, y: SN.score_forward("NCSNv2Deepest", P2, x, y)
    score_T = lambda p, y: SN.score_forward_3d_shallow(P3, p.reshape(-1, 1, 8, 8, T), y).reshape(-1, 64, T)
    prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, mask, a, l)
    adj = lambda s: M.sense_adjoint(s, maps)
    for tag, shift in (("fixed", False),):
        torch.manual_seed(304)
        np.random.seed(11)
        with torch.no_grad():
            out = ALD.ald_2dtime(score, meas, sig, 1, 1e-4, 1e4, adj, prox, mode_T="diffusion1d", lamda_T=0.5,
                                 score_T=score_T, sigmas_T=sig_T, win=8, random_shift=shift)
        assert rel_l2(out, G[f"cine_diffusion_{tag}"]) < 1e-4, tag
    # MAP baseline with the learned temporal prior (label 1 of the temporal net's OWN schedule)
    P3m = _net("NCSN3DShallow_ngf128", 13, ALD.geometric_sigmas(0.2, 0.01, 6))
    score_Tm = lambda p, y: SN.score_forward_3d_shallow(P3m, p.reshape(-1, 1, 8, 8, T), y).reshape(-1, 64, T)
    fwd = lambda x: M.sense_forward(x, maps, mask)
    x0 = adj(meas.reshape(4, T, 1, n, n)).reshape(1, T, 1, n, n)
    with torch.no_grad():
        rec = ALD.map_2dtime_tv(score, x0, meas, fwd, adj, 5e-3, 2, 1.0, 0.7, 0.3, score_T=score_Tm, win=8)
    assert rel_l2(rec, G["map2dt_diffusion"]) < 1e-4


def test_state_dict_census():
    S = _specs()
    assert len(S["NCSNv2Deepest_acdc"]) == 230 and len(S["NCSNv2_mnist28"]) == 154
    n = sum(int(np.prod(s)) for k, s in S["NCSNv2Deepest_acdc"] if k != "sigmas")
    assert abs(n - 94.13e6) < 0.02e6


def test_sigmas():
    s = ALD.geometric_sigmas(348, 0.01, 2311)
    assert s.dtype == torch.float32 and len(s) == 2311
    assert abs(float(s[0]) - 348) < 1e-3 and abs(float(s[-1]) - 0.01) < 1e-8


def test_samplers(golden):
    G = golden("samplers")
    with torch.no_grad():
        # cfg 1 shaped
        sig = ALD.geometric_sigmas(20.0, 0.01, 10)
        P = _net("NCSNv2_ngf8_28", 3, sig)
        score = lambda x, y: SN.score_forward("NCSNv2", P, x, y)
        torch.manual_seed(101)
        x0 = torch.rand(2, 1, 28, 28)
        out = ALD.ald_unconditional(score, x0, sig, 2, 6.2e-6)
        assert rel_l2(out, G["uncond_final"]) < 1e-5
        # sde 'ald' corrector
        torch.manual_seed(404)
        x = torch.rand(2, 1, 28, 28)
        t = torch.tensor([0.6, 0.2])
        std = 0.01 * (20.0 / 0.01) ** t
        sfn = lambda x, t: score(x, torch.round((1 - t) * 9).long())
        xo, xm = ALD.sde_ald_corrector(sfn, x, t, std, 0.176, 3)
        assert rel_l2(xo, G["sde_x"]) < 1e-5 and rel_l2(xm, G["sde_mean"]) < 1e-5
        # cfg 2 shaped
        n = 32
        sig = ALD.geometric_sigmas(30.0, 0.01, 10)
        P = _net("NCSNv2Deepest_ngf8", 4, sig)
        score = lambda x, y: SN.score_forward("NCSNv2Deepest", P, x, y)
        maps = M.exp_coil_maps(4, n, n, 0)
        kc = M.keep_center_mask(n, 4, 1 / 8, seed=0)
        meas = M.sense_forward(phantom(1401, 1, 1, n, n), maps, kc).repeat(1, 2, 1, 1, 1)
        adj = lambda s: M.sense_adjoint(s, maps)
        for tag, lr_scaled in (("lr1e6", 1e6), ("lr1", 1.0)):
            prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, kc, a, l)
            torch.manual_seed(202)
            out = ALD.ald_sense_real_imag(score, meas, sig, 2, 9e-7, lr_scaled, adj, prox)
            assert rel_l2(out, G[f"sense_final_{tag}"]) < 2e-5, tag
        # cfg 4 shaped
        sig = ALD.geometric_sigmas(20.0, 0.01, 10)
        P = _net("NCSNv2Deepest_ngf8", 5, sig)
        score = lambda x, y: SN.score_forward("NCSNv2Deepest", P, x, y)
        maps = M.exp_coil_maps(4, n, n, 0)
        mask = M.live_sense_mask(n, 0)
        meas = M.sense_forward(phantom(1402, 24, 1, n, n), maps, mask).reshape(4, 1, 24, 1, n, n)
        for mode_T in ("none", "tv"):
            prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, mask, a, l)
            torch.manual_seed(303)
            out = ALD.ald_2dtime(score, meas, sig, 1, 1e-4, 1e4, adj, prox, mode_T=mode_T, lamda_T=0.05)
            assert rel_l2(out, G[f"cine_final_{mode_T}"]) < 2e-5, mode_T


def test_posterior_stats_and_metrics():
    x = crandn(5, 7, 1, 16, 16)
    st = ALD.posterior_stats(x)
    xn = x.numpy()
    assert np.allclose(st["mag_mean"].numpy(), np.abs(xn).mean(0), atol=1e-6)
    assert np.allclose(st["mag_std"].numpy(), np.abs(xn).std(0), atol=1e-6)
    assert np.allclose(st["phase_std"].numpy(), np.angle(xn).std(0), atol=1e-5)
    a = rrand(1, 32, 32)
    assert abs(ALD.ssim(a, a, data_range=1.0) - 1.0) < 1e-12
    assert ALD.nrmse(a, a) == 0.0
    b = a + 0.1 * rrandn(2, 32, 32)
    assert 0 < ALD.ssim(b, a, data_range=1.0) < 1 and ALD.nrmse(b, a) > 0


def test_map_baselines(golden):
    G = golden("map")
    n = 32
    with torch.no_grad():
        sig = ALD.geometric_sigmas(30.0, 0.01, 10)
        P = _net("NCSNv2Deepest_ngf8", 6, sig)
        score = lambda x, y: SN.score_forward("NCSNv2Deepest", P, x, y)
        maps = M.exp_coil_maps(4, n, n, 0)
        kc = M.keep_center_mask(n, 4, 1 / 8, seed=0)
        fwd = lambda v: M.sense_forward(v, maps, kc)
        adj = lambda s: M.sense_adjoint(s, maps)
        y = fwd(phantom(1501, 1, 1, n, n))
        out = ALD.map_sense(score, adj(y), y, fwd, adj, 0.5, 1e-2, 6)
        assert rel_l2(out, G["map2d_final"]) < 1e-5
        sig = ALD.geometric_sigmas(20.0, 0.01, 10)
        P = _net("NCSNv2Deepest_ngf8", 5, sig)
        score = lambda x, y: SN.score_forward("NCSNv2Deepest", P, x, y)
        mask = M.live_sense_mask(n, 0)
        fwd = lambda v: M.sense_forward(v, maps, mask)
        y6 = fwd(phantom(1402, 24, 1, n, n)).reshape(4, 1, 24, 1, n, n)
        x0 = adj(y6.reshape(4, 24, 1, n, n)).reshape(1, 24, 1, n, n)
        out = ALD.map_2dtime_tv(score, x0, y6, fwd, adj, 5e-3, 3, 1.0, 0.7, 0.05)
        assert rel_l2(out, G["map2dt_final"]) < 1e-5


def test_deeper_langevin_seg_guidance(golden):
    G = golden("extra")
    with torch.no_grad():
        sig = ALD.geometric_sigmas(30.0, 0.01, 12)
        P = _net("NCSNv2Deeper_ngf8", 8, sig)
        out = SN.score_forward("NCSNv2Deeper", P, rrand(1303, 2, 1, 32, 32), torch.tensor([3, 10]))
        assert rel_l2(out, G["deeper_out"]) < 1e-5
        sig = ALD.geometric_sigmas(20.0, 0.01, 10)
        P = _net("NCSNv2_ngf8_28", 3, sig)
        score = lambda x, y: SN.score_forward("NCSNv2", P, x, y)
        sfn = lambda x, t: score(x, torch.round((1 - t) * 9).long())
        torch.manual_seed(405)
        x = torch.rand(2, 1, 28, 28)
        xo, xm = ALD.sde_langevin_corrector(sfn, x, torch.tensor([0.6, 0.2]), 0.16, 2)
        assert rel_l2(xo, G["lang_x"]) < 1e-5 and rel_l2(xm, G["lang_mean"]) < 1e-5
    # segmentation-guided chain
    n, B = 32, 2
    sig = ALD.geometric_sigmas(30.0, 0.01, 10)
    P = _net("NCSNv2Deepest_ngf8", 4, sig)
    score = lambda x, y: SN.score_forward("NCSNv2Deepest", P, x, y)
    maps = M.exp_coil_maps(4, n, n, 0)
    kc = M.keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = M.sense_forward(phantom(1401, 1, 1, n, n), maps, kc).repeat(1, B, 1, 1, 1)
    torch.manual_seed(9)
    seg = torch.nn.Conv2d(1, 3, 3, padding=1)
    label = (rrand(1601, B, 1, n, n) * 3).long().clamp(max=2)
    w = torch.linspace(0, 1, 10)                       # get_lh_weights(sigmas, 0.0, "linear"), ALD_optimizers.py:23-38
    guide = lambda xp, c: ALD.seg_guidance_grad(seg, xp, label) / sig[c] * w[c]
    prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, kc, a, l)
    torch.manual_seed(203)
    with torch.no_grad():
        out = ALD.ald_sense_real_imag(score, meas, sig, 2, 9e-7, 1e6, lambda s: M.sense_adjoint(s, maps), prox, guide=guide)
    assert rel_l2(out, G["seg_final"]) < 2e-5
