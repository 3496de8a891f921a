"""GPU suite (-m gpu): the CUDA path through the C ABI against the golden fixtures / the oracle."""
import ctypes

import numpy as np
import pytest
import torch

import parity_cases as C
from conftest import rel_l2
from oracle import mri_ops as M, scorenet as SN, ald as OALD
from oracle.fixture_inputs import crandn, rrand, rrandn, phantom

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib():
    from inverseproblemwithdiffusionmodel_b200 import _lib
    return _lib


def test_library_is_native():
    L = _lib()
    assert L.lib().ipdm_abi_version() == 3
    before = L.lib().ipdm_launch_count()
    C.i2k_complex(torch.zeros(1, 1, 8, 8, dtype=torch.complex64, device=DEV))
    torch.cuda.synchronize()
    assert L.lib().ipdm_launch_count() == before + 2


def test_fft():
    C.case_fft(DEV)


@pytest.mark.parametrize("H,W", [(8, 8), (16, 64), (128, 128), (256, 256), (512, 512), (64, 512), (512, 32)])
def test_fft_sizes_vs_oracle(H, W):
    x = crandn(H * 1000 + W, 3, 1, H, W)
    assert rel_l2(C.i2k_complex(x.to(DEV)).cpu(), M.i2k(x)) < 1e-5
    assert rel_l2(C.k2i_complex(x.to(DEV)).cpu(), M.k2i(x)) < 1e-5
    # round trip
    assert rel_l2(C.k2i_complex(C.i2k_complex(x.to(DEV))).cpu(), x) < 1e-5


def test_fft_rejects_unsupported():
    L = _lib()
    with pytest.raises(L.IpdmError):
        C.i2k_complex(torch.zeros(1, 1, 28, 28, dtype=torch.complex64, device=DEV))


def test_sense_and_prox():
    C.case_prox(DEV)
    C.case_tv(DEV)


@pytest.mark.parametrize("n,nc,B,R", [(128, 4, 3, 4), (256, 4, 2, 40), (256, 8, 1, 16), (512, 4, 1, 8)])
def test_sense_full_size(n, nc, B, R):
    """cfg-2 / cfg-5 sized operator: vs the oracle, plus the adjoint dot-product identity on masked data."""
    A = C.SENSE("exp", nc, R, 1 / 64, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, R, 1 / 64, seed=0)
    x = crandn(n + nc, B, 1, n, n)
    S = A(x.to(DEV))
    Sref = M.sense_forward(x, A.sens_maps, A.random_under_fourier.mask)
    assert rel_l2(S.cpu(), Sref) < 1e-5
    y = (A.random_under_fourier.mask * crandn(n + nc + 1, nc, B, 1, n, n)).to(torch.complex64)
    adj = A.conj_op(y.to(DEV))
    assert rel_l2(adj.cpu(), M.sense_adjoint(y, A.sens_maps)) < 1e-5
    assert rel_l2(A.conj_op_masked(y.to(DEV)).cpu(), adj.cpu()) < 1e-6
    lhs = torch.vdot(y.reshape(-1).to(DEV), S.reshape(-1))
    rhs = torch.vdot(adj.reshape(-1), x.reshape(-1).to(DEV))
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    # linearity
    x2 = crandn(n + nc + 2, B, 1, n, n).to(DEV)
    assert rel_l2((A(x.to(DEV) + 2 * x2)).cpu(), (S + 2 * A(x2)).cpu()) < 1e-5


def _sparse_mask(g, frames, W, lines):
    """`lines` sampled columns per frame: a centre window plus scattered ones (frame f keeps lines - (f % 3))."""
    mask = torch.zeros(frames, 1, 1, W, dtype=torch.bool)
    for f in range(frames):
        n = max(2, lines - (f % 3))
        win = max(2, n // 3)
        mask[f, ..., W // 2 - win // 2:W // 2 - win // 2 + win] = True
        while int(mask[f].sum()) < n:
            mask[f, ..., int(torch.randint(0, W, (1,), generator=g))] = True
    return mask


def _sense_battery(H, W, nc, B, frames, cplx, lines, use_plan):
    """Forward, unmasked / masked adjoint, dirty input, SSOS and the fused step through the C ABI against the oracle:
    mask given as a raw device mask (general kernels) or as a compiled plan (pruned kernels when it allows)."""
    L = _lib()
    lib = L.lib()
    g = torch.Generator().manual_seed(H + 3 * W + nc + (lines or 0))
    maps = torch.rand(nc, H, W, generator=g, dtype=torch.float64) + 0.2
    if cplx:
        maps = maps * torch.exp(1j * torch.rand(nc, H, W, generator=g, dtype=torch.float64))
    if lines is None:
        mask = torch.rand(frames, 1, 1, W, generator=g) < 0.15
        mask[..., W // 2 - 2:W // 2 + 2] = True
    else:
        mask = _sparse_mask(g, frames, W, lines)
    m8h = mask.reshape(frames, W).to(torch.uint8).contiguous()
    m8 = m8h.to(DEV)
    plan = L.SensePlan(m8h.numpy(), H, W) if use_plan else None
    if use_plan and lines is not None:
        # pruned unless some residue class (column mod 16) of some frame holds more than 4 sampled columns
        cmax = max(int(torch.bincount(torch.nonzero(m8h[f])[:, 0] % 16, minlength=16).max()) for f in range(frames))
        assert plan.ns_max == lines and plan.pruned == (cmax <= 4), (plan.pruned, cmax)
    mask = mask[0] if frames == 1 else mask.repeat(B // frames, 1, 1, 1)     # kernel: frame = b % frames
    x = crandn(H + W, B, 1, H, W)
    mre = maps.real.float().contiguous().to(DEV)
    mim = maps.imag.float().contiguous().to(DEV) if cplx else None
    ws = torch.empty(lib.ipdm_sense_workspace_bytes(nc, B, H, W), dtype=torch.uint8, device=DEV)
    st = L.stream()

    def fwd(xd, out):
        if use_plan:
            L.check(lib.ipdm_sense_forward_plan(plan.handle, xd.data_ptr(), mre.data_ptr(), L.ptr(mim), out.data_ptr(), nc, B, ws.data_ptr(), st), "fwd")
        else:
            L.check(lib.ipdm_sense_forward(xd.data_ptr(), mre.data_ptr(), L.ptr(mim), m8.data_ptr(), frames, out.data_ptr(), nc, B, H, W, ws.data_ptr(), st), "fwd")

    def adj_masked(yd, out, ssos=0):
        if use_plan:
            L.check(lib.ipdm_sense_adjoint_plan(plan.handle, yd.data_ptr(), L.ptr(None if ssos else mre), L.ptr(None if ssos else mim), out.data_ptr(),
                                                nc, B, ssos, ws.data_ptr(), st), "adj")
        else:
            L.check(lib.ipdm_sense_adjoint(yd.data_ptr(), L.ptr(None if ssos else mre), L.ptr(None if ssos else mim), m8.data_ptr(), frames,
                                           out.data_ptr(), nc, B, H, W, ssos, ws.data_ptr(), st), "adj")

    S = torch.full((nc, B, 1, H, W), float("nan"), dtype=torch.complex64, device=DEV)
    xd = x.to(DEV).contiguous()
    fwd(xd, S)
    Sref = M.sense_forward(x, maps, mask)
    assert rel_l2(S.cpu(), Sref) < 1e-5
    assert float((S.cpu() * (~mask)).abs().max()) == 0.0          # every unsampled column is written, with exact zeros
    y = (mask * crandn(H + W + 1, nc, B, 1, H, W)).to(torch.complex64)
    yd = y.to(DEV).contiguous()
    ref_adj = M.sense_adjoint(y, maps.to(torch.complex64) if cplx else maps)
    out = torch.full((B, 1, H, W), float("nan"), dtype=torch.complex64, device=DEV)
    L.check(lib.ipdm_sense_adjoint(yd.data_ptr(), mre.data_ptr(), L.ptr(mim), None, 1, out.data_ptr(), nc, B, H, W, 0, ws.data_ptr(), st), "adj")
    assert rel_l2(out.cpu(), ref_adj) < 1e-5
    out.fill_(float("nan"))
    adj_masked(yd, out)
    assert rel_l2(out.cpu(), ref_adj) < 1e-5
    # an input that is NOT zero off the mask: the masked adjoint must ignore the unsampled columns
    dirty = crandn(H + W + 2, nc, B, 1, H, W).to(DEV)
    out2 = torch.empty((B, 1, H, W), dtype=torch.complex64, device=DEV)
    adj_masked(dirty, out2)
    assert rel_l2(out2.cpu(), M.sense_adjoint((mask * dirty.cpu()).to(torch.complex64), maps.to(torch.complex64) if cplx else maps)) < 1e-5
    ss = torch.empty((B, 1, H, W), dtype=torch.float32, device=DEV)
    L.check(lib.ipdm_sense_adjoint(yd.data_ptr(), None, None, None, 1, ss.data_ptr(), nc, B, H, W, 1, ws.data_ptr(), st), "ssos")
    assert rel_l2(ss.cpu(), M.sense_ssos(y)) < 1e-5
    ss.fill_(float("nan"))
    adj_masked(yd, ss, ssos=1)
    assert rel_l2(ss.cpu(), M.sense_ssos(y)) < 1e-5
    # fused Langevin + data-consistency step
    gr, nz = crandn(7, B, 1, H, W), crandn(8, B, 1, H, W)
    step, kappa = 0.21, 0.6
    nsc = float(torch.sqrt(torch.tensor(step) * 2))
    z = x + step * gr + nsc * nz
    mc = maps.to(torch.complex64) if cplx else maps
    ref = z - kappa * (M.sense_adjoint((mask * M.sense_forward(z, maps, mask)).to(torch.complex64), mc) - ref_adj)
    planar = lambda c: torch.stack([c.real.reshape(B, H, W), c.imag.reshape(B, H, W)]).contiguous().to(DEV)
    state, grad, noise, bvec = planar(x), planar(gr), planar(nz), planar(ref_adj)
    sc = L.AldScalars(step, nsc, kappa, 0.0)
    if use_plan:
        L.check(lib.ipdm_ald_sense_step_plan(plan.handle, state.data_ptr(), grad.data_ptr(), noise.data_ptr(), bvec.data_ptr(), mre.data_ptr(),
                                             L.ptr(mim), nc, B, H, sc, None, None, None, st), "ald_sense_step_plan")
    else:
        L.check(lib.ipdm_ald_sense_step(state.data_ptr(), grad.data_ptr(), noise.data_ptr(), bvec.data_ptr(), mre.data_ptr(), L.ptr(mim),
                                        m8.data_ptr(), frames, nc, B, H, W, sc, None, None, None, st), "ald_sense_step")
    got = torch.complex(state[0], state[1]).cpu().reshape(B, 1, H, W)
    assert rel_l2(got, ref) < 1e-5
    assert rel_l2(got - z, ref - z) < 1e-4
    # in-kernel noise: the pruned and the general kernels draw the same stream for the same (seed, chain, pixel, step)
    if use_plan:
        s1, s2 = planar(x), planar(x)
        ids = torch.arange(100, 100 + B, dtype=torch.int32, device=DEV)
        L.check(lib.ipdm_ald_sense_step_plan(plan.handle, s1.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), mre.data_ptr(),
                                             L.ptr(mim), nc, B, H, sc, None, None, L.rng(11, 3, ids), st), "ald_sense_step_plan")
        L.check(lib.ipdm_ald_sense_step(s2.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), mre.data_ptr(), L.ptr(mim),
                                        m8.data_ptr(), frames, nc, B, H, W, sc, None, None, L.rng(11, 3, ids), st), "ald_sense_step")
        assert rel_l2(s1.cpu(), s2.cpu()) < 1e-5
        assert rel_l2(s1.cpu(), state.cpu()) > 1e-2      # and it is noise, not the injected tensor


@pytest.mark.parametrize("H,W,nc,B,frames,cplx", [(128, 128, 4, 24, 24, False), (64, 256, 3, 5, 1, True), (512, 128, 2, 2, 1, False),
                                                  (256, 64, 5, 6, 3, True), (512, 512, 3, 2, 1, True)])
def test_sense_two_pass_engine_variants(H, W, nc, B, frames, cplx):
    """The 64..512 kernels (csrc/sense_fast.cuh) through the C ABI: per-frame masks, complex coil maps, non-square
    images, SSOS, unmasked / masked adjoint and the fused step, each against the oracle."""
    _sense_battery(H, W, nc, B, frames, cplx, None, use_plan=False)


@pytest.mark.parametrize("H,W,nc,B,frames,cplx,lines", [
    (128, 128, 4, 24, 24, False, 8), (128, 128, 2, 3, 1, True, 16), (64, 256, 3, 5, 1, True, 11), (512, 128, 2, 2, 1, False, 9),
    (256, 256, 4, 6, 3, True, 20), (256, 256, 4, 14, 1, False, 11), (256, 256, 2, 2, 1, False, 32), (512, 512, 3, 2, 1, True, 21),
    (256, 512, 16, 4, 1, False, 21), (512, 512, 32, 4, 1, False, 32), (64, 512, 2, 3, 1, False, 2)])
def test_sense_pruned_plan_variants(H, W, nc, B, frames, cplx, lines):
    """The pruned kernels (csrc/sense_pruned.cuh, masks that keep <= 32 columns) through the plan entry points: one and two
    outputs per thread, every row length, per-frame masks with different line counts, complex maps, non-square images,
    16 and 32 coils at 512 columns (cfg 5), each against the oracle."""
    _sense_battery(H, W, nc, B, frames, cplx, lines, use_plan=True)


def test_sense_pruned_big_batch_overlaps_on_a_side_stream():
    """k-space >= 512 MB with >= 16 images and ipdm_debug_option(5, 1) (off by default: measured slower): the plan entry
    points split the batch into four image sub-ranges and run the column kernel of each on the plan's side stream (fork /
    join by events).  Same numbers as the general kernels, also when the call is captured into a CUDA graph and replayed."""
    L = _lib()
    lib = L.lib()
    L.check(lib.ipdm_debug_option(5, 1), "split on")
    try:
        _split_stream_body(L, lib)
    finally:
        L.check(lib.ipdm_debug_option(5, 0), "split off")


def _split_stream_body(L, lib):
    nc, B, n = 16, 16, 512
    A = C.SENSE("exp", nc, 40, 1 / 64, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 64, seed=0)
    dev = torch.device(DEV, 0)
    plan = A.device_plan(dev, n)
    assert plan.pruned
    m8, frames = A.device_mask(dev)
    mre, _ = A.device_maps(dev)
    g = torch.Generator(device=DEV).manual_seed(12)
    x = torch.randn(B, 1, n, n, generator=g, device=DEV, dtype=torch.complex64)
    ws = torch.empty(lib.ipdm_sense_workspace_bytes(nc, B, n, n), dtype=torch.uint8, device=DEV)
    S_plan = torch.full((nc, B, 1, n, n), float("nan"), dtype=torch.complex64, device=DEV)
    S_gen = torch.empty_like(S_plan)
    L.check(lib.ipdm_sense_forward_plan(plan.handle, x.data_ptr(), mre.data_ptr(), None, S_plan.data_ptr(), nc, B, ws.data_ptr(), L.stream()), "fwd")
    L.check(lib.ipdm_sense_forward(x.data_ptr(), mre.data_ptr(), None, m8.data_ptr(), frames, S_gen.data_ptr(), nc, B, n, n, ws.data_ptr(), L.stream()), "fwd")
    torch.cuda.synchronize()
    assert rel_l2(S_plan.cpu(), S_gen.cpu()) < 2e-6
    ref = M.sense_forward(x[[0, B - 1]].cpu(), A.sens_maps, A.random_under_fourier.mask)
    assert rel_l2(S_plan[:, [0, B - 1]].cpu(), ref) < 1e-5
    o_plan = torch.full((B, 1, n, n), float("nan"), dtype=torch.complex64, device=DEV)
    o_gen = torch.empty_like(o_plan)
    L.check(lib.ipdm_sense_adjoint_plan(plan.handle, S_gen.data_ptr(), mre.data_ptr(), None, o_plan.data_ptr(), nc, B, 0, ws.data_ptr(), L.stream()), "adj")
    L.check(lib.ipdm_sense_adjoint(S_gen.data_ptr(), mre.data_ptr(), None, m8.data_ptr(), frames, o_gen.data_ptr(), nc, B, n, n, 0, ws.data_ptr(), L.stream()), "adj")
    torch.cuda.synchronize()
    assert rel_l2(o_plan.cpu(), o_gen.cpu()) < 2e-6
    L.check(lib.ipdm_sense_adjoint_plan(plan.handle, S_plan.data_ptr(), mre.data_ptr(), None, o_plan.data_ptr(), nc, B, 0, ws.data_ptr(), L.stream()), "adj")
    torch.cuda.synchronize()
    # captured: the side stream joins the capture through the fork event and leaves it through the join event
    S_cap = torch.zeros_like(S_plan)
    o_cap = torch.zeros_like(o_plan)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        L.check(lib.ipdm_sense_forward_plan(plan.handle, x.data_ptr(), mre.data_ptr(), None, S_cap.data_ptr(), nc, B, ws.data_ptr(), L.stream()), "fwd")
        L.check(lib.ipdm_sense_adjoint_plan(plan.handle, S_cap.data_ptr(), mre.data_ptr(), None, o_cap.data_ptr(), nc, B, 0, ws.data_ptr(), L.stream()), "adj")
    for _ in range(2):
        S_cap.fill_(float("nan")); o_cap.fill_(float("nan"))
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(S_cap, S_plan) and torch.equal(o_cap, o_plan)


def test_sense_plan_falls_back_to_the_general_kernels():
    """A mask that keeps too many columns for the pruned kernels (141 of 512 at R = 4) still works through a plan."""
    _sense_battery(128, 512, 2, 2, 1, False, None, use_plan=True)
    L = _lib()
    m = (torch.rand(1, 256) < 0.3).to(torch.uint8)
    assert not L.SensePlan(m.numpy(), 256, 256).pruned
    m = torch.zeros(1, 64, dtype=torch.uint8); m[0, 30:34] = 1
    assert not L.SensePlan(m.numpy(), 64, 64).pruned          # W = 64 is served by the general kernels


def test_sense_many_images():
    """ncoils * batch beyond 65535 (the image index lives on grid.x): forward / adjoint of 20000 chains of 64x64, spot-checked."""
    n, nc, B = 64, 4, 20000
    A = C.SENSE("exp", nc, 8, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 8, 1 / 8, seed=0)
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.complex(torch.randn(B, 1, n, n, generator=g, device=DEV), torch.randn(B, 1, n, n, generator=g, device=DEV))
    S = A(x)
    back = A.conj_op_masked(S)
    idx = [0, 1, 7777, 16383, 16384, B - 1]
    xs = x[idx].cpu()
    Sref = M.sense_forward(xs, A.sens_maps, A.random_under_fourier.mask)
    assert rel_l2(S[:, idx].cpu(), Sref) < 1e-5
    assert rel_l2(back[idx].cpu(), M.sense_adjoint(Sref, A.sens_maps)) < 1e-5


def _conv_pair(N, H, W, Cin, Cout, taps, dil, flags, bias, residual, want32, want16, stats):
    """Run igemm and direct kernels on the same operands; return both outputs."""
    L = _lib()
    g = torch.Generator().manual_seed(H * 7 + Cin + Cout + taps + dil + flags)
    x16 = (torch.randn(N, H, W, Cin, generator=g)).half().to(DEV)
    w16 = (torch.randn(Cout, taps, Cin, generator=g) / (taps * Cin) ** 0.5).half().to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV) if bias else None
    pool = bool(flags & 8)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    res = torch.randn(N, Ho, Wo, Cout, generator=g).to(DEV) if residual else None
    outs = []
    for fn in (L.lib().ipdm_conv_igemm, L.lib().ipdm_conv_direct):
        o32 = torch.full((N, Ho, Wo, Cout), float("nan"), device=DEV) if want32 else None
        o16 = torch.full((N, Ho, Wo, Cout), float("nan"), device=DEV, dtype=torch.float16) if want16 else None
        st = torch.zeros(N, Cout, 2, device=DEV, dtype=torch.float64) if stats else None
        d = L.ConvDesc(x16.data_ptr(), w16.data_ptr(), L.ptr(b), L.ptr(res), L.ptr(o32), L.ptr(o16), L.ptr(st),
                       N, H, W, Cin, Cout, taps, dil, flags)
        L.check(fn(ctypes.byref(d), L.stream()), "conv")
        torch.cuda.synchronize()
        outs.append((o32, o16, st))
    return outs


@pytest.mark.parametrize("N,H,W,Cin,Cout,taps,dil,flags,bias,residual", [
    (1, 16, 16, 64, 128, 9, 1, 0, False, False),          # single tile, single k-chunk per tap
    (2, 32, 32, 128, 128, 9, 1, 0, True, True),
    (1, 32, 32, 256, 512, 9, 2, 0, True, False),          # dilation 2, 4 channel tiles
    (1, 32, 32, 512, 256, 9, 4, 1, False, True),          # dilation 4, ELU'd f16
    (1, 64, 64, 128, 256, 1, 1, 8, True, False),          # 1x1 + pool (ConvMeanPool shortcut)
    (2, 64, 64, 128, 256, 9, 1, 8 | 1, True, True),       # 3x3 + pool + residual
    (1, 14, 14, 256, 256, 9, 2, 4 | 2, False, True),      # partial tile (MNIST 14x14), CRP flags
    (1, 40, 24, 128, 128, 9, 1, 1, True, True),           # ragged tiles both ways
    (3, 128, 128, 128, 128, 9, 1, 0, False, True),        # many tiles
    (4, 256, 256, 128, 128, 9, 1, 1, False, True),        # more work items than SMs: persistent loop, both accumulators
    (2, 64, 64, 256, 256, 9, 2, 0, True, True),           # dilation 2 in the halo kernel
])
@pytest.mark.parametrize("variant", [0, 1])      # 0: auto (persistent halo kernel where it applies), 1: per-tap tile kernel
def test_conv_igemm_vs_direct(N, H, W, Cin, Cout, taps, dil, flags, bias, residual, variant):
    L = _lib()
    L.check(L.lib().ipdm_debug_option(1, variant))
    try:
        (a32, a16, ast), (b32, b16, bst) = _conv_pair(N, H, W, Cin, Cout, taps, dil, flags, bias, residual, True, True, True)
    finally:
        L.check(L.lib().ipdm_debug_option(1, 0))
    assert not torch.isnan(a32).any() and not torch.isnan(a16.float()).any()
    assert rel_l2(a32.cpu(), b32.cpu()) < 1e-5      # fp32 accumulation order over K = taps*Cin differs
    assert rel_l2(a16.float().cpu(), b16.float().cpu()) < 1e-3       # f16 rounding-boundary flips only
    HW = a32.shape[1] * a32.shape[2]
    ref_st = torch.stack([b32.reshape(N, HW, Cout).sum(1), (b32 ** 2).reshape(N, HW, Cout).sum(1)], dim=-1)
    assert rel_l2(ast.float().cpu(), ref_st.cpu()) < 1e-5
    assert rel_l2(bst.float().cpu(), ref_st.cpu()) < 1e-5


@pytest.mark.parametrize("want32,want16,stats", [(True, False, True), (False, True, False), (True, False, False)])
@pytest.mark.parametrize("N,H,W,Cin,Cout,dil,flags,bias,residual", [
    (2, 64, 64, 128, 128, 1, 1, True, True), (1, 40, 24, 128, 256, 2, 0, False, False), (1, 64, 32, 128, 128, 1, 8 | 1, True, True),
    (3, 96, 72, 64, 128, 1, 4 | 2 | 1, False, True),
    (4, 12, 16, 128, 128, 1, 1, True, True),      # 12-row tile, two images per work item
    (3, 12, 8, 128, 256, 2, 0, False, True),      # 12-row tile, odd image count: one image per item
    (6, 24, 8, 128, 128, 2, 1, True, True),       # 24-row tile (N = 192)
    (2, 12, 8, 256, 256, 4, 1, True, False)])     # dilation 4 on the halo kernel (12-row tile only)
def test_conv_halo_output_modes(N, H, W, Cin, Cout, dil, flags, bias, residual, want32, want16, stats):
    """Every epilogue specialisation of the persistent halo kernel (fp32-only, f16-only, with / without residual,
    pooled) against the CUDA-core direct kernel, including tiles that overhang the image."""
    (a32, a16, ast), (b32, b16, bst) = _conv_pair(N, H, W, Cin, Cout, 9, dil, flags, bias, residual, want32, want16, stats)
    if want32:
        assert not torch.isnan(a32).any()
        assert rel_l2(a32.cpu(), b32.cpu()) < 1e-5
    if want16:
        assert not torch.isnan(a16.float()).any()
        assert rel_l2(a16.float().cpu(), b16.float().cpu()) < 1e-3
    if stats:
        assert rel_l2(ast.float().cpu(), bst.float().cpu()) < 1e-5


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("N,H,W,Cin,Cout,taps,dil,flags,bias,residual,want16,stats", [
    (2, 64, 64, 128, 128, 9, 1, 1, False, True, True, False),        # RCU second convolution: residual + result + ELU copy
    (2, 64, 64, 128, 128, 9, 1, 0, True, True, False, True),         # last refine block: residual + result + sums
    (1, 40, 24, 128, 256, 9, 2, 0, True, False, False, True),        # encoder conv1: result + sums, ragged tiles
    (2, 32, 32, 256, 256, 9, 1, 4 | 2, False, True, True, False),    # CRP entry: ELU'd residual, pre-residual f16 copy
    (2, 64, 64, 128, 256, 9, 1, 8 | 1, True, True, True, True),      # pooled conv2 of a down block
    (1, 64, 64, 128, 256, 1, 1, 8, True, False, False, False),       # pooled 1x1 shortcut (per-tap kernel either way)
    (3, 24, 8, 128, 128, 9, 2, 1, False, True, True, False),         # 24-row tile
    (4, 12, 16, 128, 128, 9, 1, 1, True, True, True, False),         # 12-row tile, two images per work item
    (4, 256, 256, 128, 128, 9, 1, 1, False, True, True, False),      # the dominant launch of the benchmark
    (2, 32, 32, 256, 512, 9, 4, 0, True, True, False, False)])       # dilation 4 (per-tap kernel)
def test_conv_16bit_residual_stream(N, H, W, Cin, Cout, taps, dil, flags, bias, residual, want16, stats, variant):
    """ipdm_conv_desc.residual_f16 / out_raw_f16: the residual stream kept in 16 bits.  Same operands through the f32-stream
    path with the f16 residual widened: the result must be that path's result rounded to f16 (fp32 arithmetic in between is
    identical), the f16 operand copy and the InstanceNorm++ sums (taken from the fp32 values) must agree."""
    L = _lib()
    g = torch.Generator().manual_seed(H * 11 + Cin + Cout + taps + dil + flags)
    x16 = torch.randn(N, H, W, Cin, generator=g).half().to(DEV)
    w16 = (torch.randn(Cout, taps, Cin, generator=g) / (taps * Cin) ** 0.5).half().to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV) if bias else None
    pool = bool(flags & 8)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    res16 = (3 * torch.randn(N, Ho, Wo, Cout, generator=g)).half().to(DEV) if residual else None
    res32 = res16.float() if residual else None
    L.check(L.lib().ipdm_debug_option(1, variant))
    try:
        outs = []
        for t16 in (False, True):
            o_res = torch.full((N, Ho, Wo, Cout), float("nan"), device=DEV, dtype=torch.float16 if t16 else torch.float32)
            o16 = torch.full((N, Ho, Wo, Cout), float("nan"), device=DEV, dtype=torch.float16) if want16 else None
            st = torch.zeros(N, Cout, 2, device=DEV, dtype=torch.float64) if stats else None
            if t16:
                d = L.ConvDesc(x16.data_ptr(), w16.data_ptr(), L.ptr(b), None, None, L.ptr(o16), L.ptr(st), N, H, W, Cin, Cout, taps, dil, flags,
                               0, 0, L.ptr(res16), o_res.data_ptr())
            else:
                d = L.ConvDesc(x16.data_ptr(), w16.data_ptr(), L.ptr(b), L.ptr(res32), o_res.data_ptr(), L.ptr(o16), L.ptr(st), N, H, W, Cin, Cout,
                               taps, dil, flags)
            L.check(L.lib().ipdm_conv_igemm(ctypes.byref(d), L.stream()), "conv")
            torch.cuda.synchronize()
            outs.append((o_res, o16, st))
    finally:
        L.check(L.lib().ipdm_debug_option(1, 0))
    (a32, a16, ast), (t_raw, t16c, tst) = outs
    assert not torch.isnan(t_raw.float()).any()
    assert torch.equal(t_raw, a32.half())                      # the 16-bit stream is the fp32 result, rounded once
    if want16:
        assert torch.equal(t16c, a16)
    if stats:
        assert rel_l2(tst.cpu(), ast.cpu()) < 1e-12
    # mixing the two streams is refused
    bad = L.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, L.ptr(res32), None, None, None, N, H, W, Cin, Cout, taps, dil, flags, 0, 0, None, t_raw.data_ptr())
    if residual:
        assert L.lib().ipdm_conv_igemm(ctypes.byref(bad), L.stream()) != 0


@pytest.mark.parametrize("N,H,W,C", [(2, 14, 14, 256), (1, 28, 28, 128), (3, 40, 24, 64), (2, 256, 256, 128), (5, 32, 32, 512), (1, 7, 5, 8)])
def test_maxpool5_bit_exact(N, H, W, C):
    """5x5 / stride 1 / pad 2 max-pool of the CRP blocks (layers.py:69-80): exact against torch on the same f16 values."""
    L = _lib()
    g = torch.Generator().manual_seed(H * 31 + W)
    x = torch.randn(N, H, W, C, generator=g).half().to(DEV)
    out = torch.full_like(x, float("nan"))
    L.check(L.lib().ipdm_maxpool5_f16(x.data_ptr(), out.data_ptr(), N, H, W, C, L.stream()), "maxpool5")
    ref = torch.nn.functional.max_pool2d(x.float().permute(0, 3, 1, 2), 5, 1, 2).permute(0, 2, 3, 1).half()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("N,h,w,H,W,C,acc", [(2, 16, 16, 32, 32, 128, 1), (1, 7, 5, 14, 10, 64, 0), (3, 64, 64, 128, 128, 128, 1), (2, 8, 8, 8, 8, 256, 1)])
def test_bilinear_add_vs_torch(N, h, w, H, W, C, acc):
    """MSF block: dst (+)= bilinear(src, align_corners=True) and the f16(ELU) copy (layers.py:165-184)."""
    L = _lib()
    g = torch.Generator().manual_seed(h * 17 + W)
    src = torch.randn(N, h, w, C, generator=g).to(DEV)
    dst0 = torch.randn(N, H, W, C, generator=g).to(DEV)
    dst = dst0.clone()
    h16 = torch.full((N, H, W, C), float("nan"), device=DEV, dtype=torch.float16)
    L.check(L.lib().ipdm_bilinear_add(src.data_ptr(), dst.data_ptr(), h16.data_ptr(), N, h, w, H, W, C, acc, L.stream()), "bilinear_add")
    up = torch.nn.functional.interpolate(src.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    ref = up + dst0 if acc else up
    assert rel_l2(dst.cpu(), ref.cpu()) < 1e-6
    assert rel_l2(h16.float().cpu(), torch.nn.functional.elu(ref).half().float().cpu()) < 1e-3


@pytest.mark.parametrize("N,h,w,H,W,C,acc", [(2, 16, 16, 32, 32, 128, 1), (1, 7, 5, 14, 10, 64, 0), (3, 64, 64, 128, 128, 128, 1), (2, 8, 8, 8, 8, 256, 1),
                                             (2, 9, 6, 18, 12, 12, 1), (1, 128, 128, 256, 256, 128, 1)])
def test_bilinear_add_f16_stream_vs_torch(N, h, w, H, W, C, acc):
    """The same on the 16-bit residual stream (ipdm_bilinear_add_f16; 8 channels per thread when C % 8 == 0, else 4): f16 in,
    fp32 interpolation and add, f16 result + f16(ELU) copy -- against torch on the same f16 inputs, one f16 rounding."""
    L = _lib()
    g = torch.Generator().manual_seed(h * 19 + W)
    src = torch.randn(N, h, w, C, generator=g).half().to(DEV)
    dst0 = torch.randn(N, H, W, C, generator=g).half().to(DEV)
    dst = dst0.clone()
    h16 = torch.full((N, H, W, C), float("nan"), device=DEV, dtype=torch.float16)
    L.check(L.lib().ipdm_bilinear_add_f16(src.data_ptr(), dst.data_ptr(), h16.data_ptr(), N, h, w, H, W, C, acc, L.stream()), "bilinear_add_f16")
    up = torch.nn.functional.interpolate(src.float().permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    ref = up + dst0.float() if acc else up
    assert torch.isfinite(dst.float()).all() and torch.isfinite(h16.float()).all()
    assert (dst.float() - ref).abs().max() <= 1e-3 * ref.abs().max() + 1e-6          # half an f16 ulp at the largest value
    assert rel_l2(dst.float().cpu(), ref.cpu()) < 5e-4
    assert rel_l2(h16.float().cpu(), torch.nn.functional.elu(ref).cpu()) < 1e-3


@pytest.mark.parametrize("N,H,W,C", [(2, 16, 24, 128), (1, 256, 256, 128), (3, 6, 10, 8)])
def test_meanpool2_f16_vs_torch(N, H, W, C):
    """Operand of the pooled 1x1 shortcut: 2x2 mean of f16 values in fp32, one rounding."""
    L = _lib()
    g = torch.Generator().manual_seed(H + W + C)
    x = (torch.randn(N, H, W, C, generator=g) * 50).half().to(DEV)
    out = torch.full((N, H // 2, W // 2, C), float("nan"), dtype=torch.float16, device=DEV)
    L.check(L.lib().ipdm_meanpool2_f16(x.data_ptr(), out.data_ptr(), N, H, W, C, L.stream()), "meanpool2_f16")
    v = x.float()
    ref = ((((v[:, ::2, ::2] + v[:, 1::2, ::2]) + v[:, ::2, 1::2]) + v[:, 1::2, 1::2]) * 0.25).half()
    assert torch.equal(out, ref)


def test_conv_direct_vs_torch():
    """Anchor of the chain igemm -> direct -> torch: the CUDA-core kernel against F.conv2d on the same f16 operands."""
    import torch.nn.functional as F
    L = _lib()
    g = torch.Generator().manual_seed(3)
    N, H, W, Cin, Cout = 2, 12, 20, 16, 24
    x16 = torch.randn(N, H, W, Cin, generator=g).half()
    w16 = (torch.randn(Cout, 9, Cin, generator=g) / 12).half()
    b = torch.randn(Cout, generator=g)
    o32 = torch.empty(N, H, W, Cout, device=DEV)
    d = L.ConvDesc(x16.to(DEV).data_ptr(), 0, 0, 0, 0, 0, 0, N, H, W, Cin, Cout, 9, 2, 0)
    xd, wd, bd = x16.to(DEV), w16.to(DEV), b.to(DEV)
    d = L.ConvDesc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), None, o32.data_ptr(), None, None, N, H, W, Cin, Cout, 9, 2, 0)
    L.check(L.lib().ipdm_conv_direct(ctypes.byref(d), L.stream()), "conv_direct")
    ref = F.conv2d(x16.float().permute(0, 3, 1, 2), w16.float().permute(0, 2, 1).reshape(Cout, Cin, 3, 3), b, padding=2, dilation=2)
    assert rel_l2(o32.cpu(), ref.permute(0, 2, 3, 1)) < 2e-6


def test_scorenet_small():
    C.case_scorenet_small(DEV)


def _full_width_net(cls, arch, size, seed, num_classes=12, sigma_begin=30.0):
    cfg = C.make_config("ACDC", 128, size, num_classes, sigma_begin, device=DEV)
    net = cls(cfg)
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    Pd = SN.synth_state_dict(spec, seed, net.sigmas.cpu())
    net.load_state_dict(Pd)
    return net.to(DEV).eval(), Pd, cfg


def test_scorenet_ngf128_tensor_core_path():
    """ngf = 128: every inner convolution goes through the tcgen05 implicit GEMM."""
    L = _lib()
    net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", 32, 21)
    x = rrand(77, 2, 1, 32, 32) * 2 - 0.5
    y = torch.tensor([0, 9])
    before = L.lib().ipdm_launch_count()
    out = net(x.to(DEV), y.to(DEV)).cpu()
    with torch.no_grad():
        ref = SN.score_forward("NCSNv2Deepest", Pd, x, y)
        SN.OPERAND_ROUND = torch.float16          # the oracle with conv operands rounded to f16 like the kernels' (fp32 accumulate)
        SN.STREAM_ROUND = torch.float16 if net._plan(2, 32, 32, torch.device(DEV, 0)).t16 else None   # ... and the 16-bit residual stream
        try:
            emu = SN.score_forward("NCSNv2Deepest", Pd, x, y)
        finally:
            SN.OPERAND_ROUND = SN.STREAM_ROUND = None
    print("ngf128 Deepest score: vs fp32 oracle %.2e, vs f16-operand oracle %.2e" % (rel_l2(out, ref), rel_l2(out, emu)))
    assert rel_l2(out, ref) < C.TOL_SCORE
    assert rel_l2(out, emu) < C.TOL_SCORE_EMU
    assert L.lib().ipdm_launch_count() - before > 150
    net2, Pd2, _ = _full_width_net(C.NCSNv2, "NCSNv2", 28, 22)
    x = rrand(78, 2, 1, 28, 28)
    out = net2(x.to(DEV), y.to(DEV)).cpu()
    with torch.no_grad():
        ref = SN.score_forward("NCSNv2", Pd2, x, y)
        SN.OPERAND_ROUND = torch.float16
        SN.STREAM_ROUND = torch.float16 if net2._plan(2, 28, 28, torch.device(DEV, 0)).t16 else None
        try:
            emu = SN.score_forward("NCSNv2", Pd2, x, y)
        finally:
            SN.OPERAND_ROUND = SN.STREAM_ROUND = None
    print("ngf128 NCSNv2 score: vs fp32 oracle %.2e, vs f16-operand oracle %.2e" % (rel_l2(out, ref), rel_l2(out, emu)))
    assert rel_l2(out, ref) < C.TOL_SCORE
    assert rel_l2(out, emu) < C.TOL_SCORE_EMU


def test_f16_range_audit_finds_clipped_activations():
    """the f16 operand path needs a range check.  The audit kernel on synthetic tensors, then on a real forward:
    an ordinary input stays far inside the range; an input at |x| ~ 3e5 drives the un-normalised decoder past 65504, the
    stores saturate (the output stays finite) and the audit reports it."""
    L = _lib()
    t = torch.zeros(1000003, dtype=torch.float16, device=DEV)
    t[12345] = -1234.0
    t[999999] = 65504.0
    t[1000002] = float("inf")
    m = torch.zeros(1, device=DEV)
    c = torch.zeros(1, dtype=torch.int64, device=DEV)
    L.check(L.lib().ipdm_f16_range_audit(t.data_ptr(), t.numel(), m.data_ptr(), c.data_ptr(), L.stream()), "audit")
    assert int(c.item()) == 2 and float(m.item()) >= 65504.0
    t[999999] = 3.0
    t[1000002] = 2.0
    m.zero_(); c.zero_()
    L.check(L.lib().ipdm_f16_range_audit(t.data_ptr(), t.numel(), m.data_ptr(), c.data_ptr(), L.stream()), "audit")
    assert int(c.item()) == 0 and float(m.item()) == 1234.0
    net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", 32, 21)
    y = torch.tensor([0, 9], device=DEV)
    out = net((rrand(77, 2, 1, 32, 32) * 2 - 0.5).to(DEV), y)
    a = net.range_audit()
    assert a["saturated_total"] == 0 and 0 < a["max_abs"] < 6e4 and bool(torch.isfinite(out).all())
    out = net((rrandn(78, 2, 1, 32, 32) * 3e5).to(DEV), y)
    a = net.range_audit()
    assert a["saturated_total"] > 0 and bool(torch.isfinite(out).all())


class _Guarded:
    """A device buffer with guard bands: the kernel gets the middle, the bands must come back untouched (compute-sanitizer is
    closed on the GPU pool, so out-of-bounds WRITES are caught this way; unwritten outputs by the NaN pre-fills of the parity
    tests)."""
    GUARD = 4096     # bytes on either side

    def __init__(self, nbytes, fill=0xA5):
        self.n = nbytes
        self.raw = torch.full((nbytes + 2 * self.GUARD,), fill, dtype=torch.uint8, device=DEV)
        self.fill = fill

    def view(self, dtype, shape):
        return self.raw[self.GUARD:self.GUARD + self.n].view(dtype).view(shape)

    def intact(self):
        lo, hi = self.raw[:self.GUARD], self.raw[self.GUARD + self.n:]
        return bool((lo == self.fill).all()) and bool((hi == self.fill).all())


@pytest.mark.parametrize("n,nc,B,lines", [(256, 4, 5, 11), (512, 3, 3, 21), (128, 4, 6, 9), (64, 2, 3, 5)])
def test_no_kernel_writes_outside_its_buffers_sense(n, nc, B, lines):
    """Plan (pruned) SENSE kernels and the fused step: output, scratch and state sit between guard bands."""
    L = _lib()
    lib = L.lib()
    g = torch.Generator().manual_seed(n + lines)
    m8h = _sparse_mask(g, 1, n, lines).reshape(1, n).to(torch.uint8).contiguous()
    plan = L.SensePlan(m8h.numpy(), n, n)
    maps = (torch.rand(nc, n, n, generator=g) + 0.2).to(DEV)
    x = crandn(n, B, 1, n, n).to(DEV)
    ws_bytes = lib.ipdm_sense_workspace_bytes(nc, B, n, n)
    gws, gS, gout, gx = _Guarded(ws_bytes), _Guarded(nc * B * n * n * 8), _Guarded(B * n * n * 8), _Guarded(2 * B * n * n * 4)
    ws, S = gws.view(torch.uint8, (ws_bytes,)), gS.view(torch.complex64, (nc, B, 1, n, n))
    out, st = gout.view(torch.complex64, (B, 1, n, n)), gx.view(torch.float32, (2, B, n, n))
    L.check(lib.ipdm_sense_forward_plan(plan.handle, x.data_ptr(), maps.data_ptr(), None, S.data_ptr(), nc, B, ws.data_ptr(), L.stream()), "fwd")
    L.check(lib.ipdm_sense_adjoint_plan(plan.handle, S.data_ptr(), maps.data_ptr(), None, out.data_ptr(), nc, B, 0, ws.data_ptr(), L.stream()), "adj")
    L.check(lib.ipdm_sense_adjoint_plan(plan.handle, S.data_ptr(), None, None, out.data_ptr(), nc, B, 1, ws.data_ptr(), L.stream()), "ssos")
    st.copy_(torch.randn(2, B, n, n, generator=g))
    grad, bvec = torch.randn(2, B, n, n, generator=g).to(DEV), torch.randn(2, B, n, n, generator=g).to(DEV)
    sc = L.AldScalars(0.1, 0.4, 0.01, 1.0)
    L.check(lib.ipdm_ald_sense_step_plan(plan.handle, st.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), maps.data_ptr(), None, nc, B, n,
                                         sc, None, None, L.rng(7, 0), L.stream()), "step")
    torch.cuda.synchronize()
    assert plan.pruned == (n >= 128) and bool(torch.isfinite(st).all())      # W = 64: general kernels behind the same entry points
    for name, gb in (("workspace", gws), ("k-space", gS), ("image", gout), ("state", gx)):
        assert gb.intact(), f"{name}: guard band overwritten"


@pytest.mark.parametrize("N,H,W,Cin,Cout,pool", [(2, 40, 24, 128, 128, 0), (1, 64, 64, 128, 256, 1), (3, 32, 32, 256, 128, 0), (2, 24, 40, 128, 128, 0)])
def test_no_kernel_writes_outside_its_buffers_stream(N, H, W, Cin, Cout, pool):
    """16-bit-stream convolution (f16 residual in, f16 result + f16 ELU copy out, optionally 2x2-pooled), norm-apply and
    bilinear kernels: every output between guard bands, images that do not fill whole pixel tiles included."""
    import ctypes
    L = _lib()
    lib = L.lib()
    g = torch.Generator().manual_seed(H * 7 + W + Cout)
    x16 = torch.randn(N, H, W, Cin, generator=g).half().to(DEV)
    w16 = (torch.randn(Cout, 9, Cin, generator=g) / (9 * Cin) ** 0.5).half().to(DEV)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    nout = N * Ho * Wo * Cout
    res16 = torch.randn(N, Ho, Wo, Cout, generator=g).half().to(DEV)
    graw, gelu, gst = _Guarded(nout * 2), _Guarded(nout * 2), _Guarded(N * Cout * 2 * 8, fill=0)
    raw, elu = graw.view(torch.float16, (N, Ho, Wo, Cout)), gelu.view(torch.float16, (N, Ho, Wo, Cout))
    stats = gst.view(torch.float64, (N, Cout, 2))
    d = L.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, None, None, elu.data_ptr(), stats.data_ptr(), N, H, W, Cin, Cout, 9, 1,
                   L.CONV_F16_ELU | (L.CONV_POOL2 if pool else 0), 0, 0, res16.data_ptr(), raw.data_ptr())
    L.check(lib.ipdm_conv_igemm(ctypes.byref(d), L.stream()), "igemm t16")
    # norm-apply on the f16 stream and the f16 bilinear accumulate, into guarded outputs as well
    gop = _Guarded(nout * 2)
    op = gop.view(torch.float16, (N, Ho, Wo, Cout))
    alpha, gamma, beta = (torch.randn(Cout, generator=g).to(DEV) for _ in range(3))
    L.check(lib.ipdm_instnorm_apply_elu_f16in(raw.data_ptr(), stats.data_ptr(), 0, alpha.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                              op.data_ptr(), N, Ho * Wo, Cout, L.stream()), "norm f16in")
    src = torch.randn(N, max(Ho // 2, 1), max(Wo // 2, 1), Cout, generator=g).half().to(DEV)
    L.check(lib.ipdm_bilinear_add_f16(src.data_ptr(), raw.data_ptr(), elu.data_ptr(), N, src.shape[1], src.shape[2], Ho, Wo, Cout, 1, L.stream()), "bilinear f16")
    torch.cuda.synchronize()
    assert bool(torch.isfinite(raw.float()).all()) and bool(torch.isfinite(op.float()).all())
    for name, gb in (("raw stream", graw), ("ELU copy", gelu), ("norm sums", gst), ("operand", gop)):
        assert gb.intact(), f"{name}: guard band overwritten"


def test_operand_shift_narrow_net():
    """Range safety: forced shift == golden score; clipped activations make the first forward escalate by itself (CUDA-core convs)."""
    C.case_operand_shift(DEV)


def test_operand_shift_tensor_core_path(monkeypatch):
    """The same on the tensor-core kernels (ngf 128): an input at |x| ~ 3e5 drives the un-normalised decoder past 65504; the first
    forward notices, moves to operand shift >= 6 + fp32 stream and then agrees with the fp32 oracle at the usual f16-operand
    level, while the clipped run (IPDM_ALLOW_F16_SATURATION=1) is finite but far off.  A forced shift on an ordinary input
    gives the unshifted score up to rounding."""
    import warnings
    y = torch.tensor([0, 9], device=DEV)
    big = rrandn(78, 2, 1, 32, 32) * 3e5
    net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", 32, 21)
    with torch.no_grad():
        ref = SN.score_forward("NCSNv2Deepest", Pd, big, y.cpu())
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        out = net(big.to(DEV), y)
    plan = next(iter(net._plans.values()))
    assert plan.shift >= 6 and not plan.t16 and any("operand exponent shift" in str(w.message) for w in wlist)
    assert net.range_audit()["saturated_total"] == 0
    e_shift = rel_l2(out.cpu(), ref)
    ordinary = (rrand(77, 2, 1, 32, 32) * 2 - 0.5)
    with torch.no_grad():
        ref_o = SN.score_forward("NCSNv2Deepest", Pd, ordinary, y.cpu())
    e_ord_shifted = rel_l2(net(ordinary.to(DEV), y).cpu(), ref_o)          # the plan stays shifted: still the right score
    net0, _, _ = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", 32, 21)
    e_ord = rel_l2(net0(ordinary.to(DEV), y).cpu(), ref_o)
    monkeypatch.setenv("IPDM_ALLOW_F16_SATURATION", "1")
    net2, _, _ = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", 32, 21)
    clipped = net2(big.to(DEV), y)
    e_clip = rel_l2(clipped.cpu(), ref)
    print(f"shift {plan.shift}: big input {e_shift:.2e} (clipped run {e_clip:.2e}); ordinary input {e_ord_shifted:.2e} shifted vs {e_ord:.2e} unshifted")
    assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(clipped).all())
    assert e_shift < 5e-3 and e_clip > 10 * e_shift
    assert e_ord < C.TOL_SCORE and e_ord_shifted < 2 * C.TOL_SCORE


def test_single_ald_step_ngf128():
    """north_star: a single cfg-2 ALD step (2 score forwards + update + prox) within 1e-4 relative L2,
    at the first level from x0 = A^H y (the hardest state) and in the steady state."""
    n, B = 64, 2
    net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", n, 23, num_classes=2311, sigma_begin=348.0)
    sig = C.get_sigmas(cfg, mode="recons")
    A = C.SENSE("exp", 4, 40, 1 / 16, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 16, seed=0)
    meas = A(phantom(11, 1, 1, n, n).to(DEV)).repeat(1, B, 1, 1, 1)
    maps, mask = A.sens_maps, A.random_under_fourier.mask
    score = lambda x, y: SN.score_forward("NCSNv2Deepest", Pd, x, y)
    prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, mask, a, l)
    adj = lambda s: M.sense_adjoint(s, maps)
    draw = lambda shape: torch.randn(*shape)
    params = {"n_steps_each": 1, "step_lr": 9e-7, "denoise": False, "final_only": True}
    for levels in (sig[:1], sig[1155:1156], sig[-1:]):
        # a two-entry schedule holding the same sigma twice: install it as the net's own sigma table so the
        # score is divided by the right sigma, and fold (sigma/sigma_L)^2 into step_lr (the loop sees ratio 1)
        lv2 = levels.repeat(2)
        net.sigmas = lv2.to(DEV)
        Pd["sigmas"] = lv2.cpu()
        ratio2 = float((levels[0] / sig[-1]) ** 2)
        sampler = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, lv2.to(DEV),
                                                  dict(params, step_lr=9e-7 * ratio2), cfg,
                                                  measurement=meas, linear_tfm=A, seg=None, device=torch.device(DEV))
        torch.manual_seed(6)
        got = sampler(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6 / ratio2, seg_mode="full", noise_fn=draw)[0]
        torch.manual_seed(6)
        with torch.no_grad():
            ref = OALD.ald_sense_real_imag(score, meas.cpu(), lv2.cpu(), 1, 9e-7 * ratio2, 1e6 / ratio2, adj, prox, denoise=False)
        torch.set_grad_enabled(True)
        assert rel_l2(got, ref) < 1e-4, float(levels[0])


def test_pooled_convs_strided_forms_tensor_core(monkeypatch):
    """ConvMeanPool as a 4x4 stride-2 convolution on space-to-depth operands with the tap mask (16 of 36 weight blocks), and the
    pooled 1x1 shortcuts on pooled operands, on the tensor-core kernels: forced on a 64x64 problem (the size heuristic would keep
    the reference's order there), against the fp32 oracle and against the pool-after-conv order."""
    y = torch.tensor([0, 9], device=DEV)
    x = rrand(77, 2, 1, 64, 64) * 2 - 0.5
    monkeypatch.setenv("IPDM_POOL_STRIDED_ALWAYS", "1")
    net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", 64, 21)
    out = net(x.to(DEV), y).cpu()
    plan = next(iter(net._plans.values()))
    assert any(k.endswith(".xp16") for k in plan.bufs) and any(k.endswith(".conv2.conv.s2d") and v[2] for k, v in plan.w.items())
    monkeypatch.delenv("IPDM_POOL_STRIDED_ALWAYS")
    net2, _, _ = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", 64, 21)
    ref_order = net2(x.to(DEV), y).cpu()
    assert not any(k.endswith(".xp16") for k in next(iter(net2._plans.values())).bufs)
    with torch.no_grad():
        ref = SN.score_forward("NCSNv2Deepest", Pd, x, y.cpu())
    e_new, e_old, e_pair = rel_l2(out, ref), rel_l2(ref_order, ref), rel_l2(out, ref_order)
    print(f"strided forms: {e_new:.2e} vs fp32 oracle (pool-after-conv: {e_old:.2e}); the two orders differ by {e_pair:.2e}")
    assert e_new < C.TOL_SCORE and e_old < C.TOL_SCORE and e_pair < C.TOL_SCORE_EMU


def test_sampler_range_check_sees_the_real_state():
    """The captured-graph fast path primes its step graph on dummy zeros, so the sampler first runs one eager score forward on the
    REAL initial state: with a measurement 3e5 times too large the plan must have moved to an operand shift before the graph
    was captured (finite chain), with an ordinary one it must not."""
    import warnings
    n, B = 32, 2
    A = C.SENSE("exp", 4, 8, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 8, 1 / 8, seed=0)
    params = {"n_steps_each": 2, "step_lr": 9e-7, "denoise": False, "final_only": True}
    for scale, want_shift in ((1.0, False), (3e5, True)):
        net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", n, 21, num_classes=4, sigma_begin=30.0)
        sig = C.get_sigmas(cfg, mode="recons")
        meas = A(phantom(14, 1, 1, n, n).to(DEV)).repeat(1, B, 1, 1, 1) * scale
        sampler = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig.to(DEV), params, cfg,
                                                  measurement=meas, linear_tfm=A, seg=None, device=torch.device(DEV))
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            out = sampler(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1.0, seg_mode="full", seed=3)[0]
        torch.set_grad_enabled(True)
        plan = next(iter(net._plans.values()))
        assert bool(torch.isfinite(out).all())
        assert (plan.shift >= 6) == want_shift, (scale, plan.shift)
        assert any("operand exponent shift" in str(w.message) for w in wlist) == want_shift
        assert net.range_audit()["saturated_total"] == 0      # the last replayed step left nothing clipped either


def test_single_ald_step_ngf128_bench_config():
    """The benchmarked configuration itself: 256x256, 14 chains (28 images per forward: work items >> SMs), 4 coils, R = 40,
    ngf 128 on the tensor-core path.  One ALD step (2 score forwards + update + prox) with injected noise at the first
    level from x0 = A^H y and at a middle level; chains are independent, so the oracle runs the first and the last
    chain only (14 chains x 2 forwards at 256^2 would take minutes of CPU).  north_star bar: 1e-4 relative L2."""
    n, B = 256, 14
    net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", n, 29, num_classes=2311, sigma_begin=348.0)
    sig = C.get_sigmas(cfg, mode="recons")
    A = C.SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 64, seed=0)
    assert A.device_plan(torch.device(DEV, 0), n).pruned
    meas = A(phantom(12, B, 1, n, n).to(DEV))                       # 14 different images
    maps, mask = A.sens_maps, A.random_under_fourier.mask
    score = lambda x, y: SN.score_forward("NCSNv2Deepest", Pd, x, y)
    prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, mask, a, l)
    adj = lambda s: M.sense_adjoint(s, maps)
    params = {"n_steps_each": 1, "step_lr": 9e-7, "denoise": False, "final_only": True}
    sub = [0, B - 1]
    for li in (0, 1155):
        lv = sig[li:li + 1]
        net.sigmas = lv.to(DEV)
        Pd["sigmas"] = lv.cpu()
        ratio2 = float((lv[0] / sig[-1]) ** 2)
        g = torch.Generator().manual_seed(600 + li)
        noises = [torch.randn(B, 1, n, n, generator=g) for _ in range(2)]     # real, then imaginary (:238-241)
        it = iter(noises)
        sampler = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, lv.to(DEV),
                                                  dict(params, step_lr=9e-7 * ratio2), cfg,
                                                  measurement=meas, linear_tfm=A, seg=None, device=torch.device(DEV))
        got = sampler(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6 / ratio2, seg_mode="full", noise_fn=lambda shape: next(it))[0]
        it2 = iter([nz[sub] for nz in noises])
        with torch.no_grad():
            ref = OALD.ald_sense_real_imag(score, meas.cpu()[:, sub], lv.cpu(), 1, 9e-7 * ratio2, 1e6 / ratio2, adj, prox, denoise=False,
                                           draw=lambda shape: next(it2))
        torch.set_grad_enabled(True)
        assert rel_l2(got[sub], ref) < 1e-4, (li, rel_l2(got[sub], ref))


def test_full_chain_metrics_ngf128_tensor_core_path():
    """north_star bar 3 on the TENSOR-CORE path: a complete chain (12 levels x 3 steps + denoise) of 2 chains at 64x64 with
    ngf 128 and identical injected noise; NRMSE / SSIM of the posterior mean and the mean of per-chain metrics within 1e-3 of
    the oracle's, and the samples themselves within 1e-3."""
    n, B, levels = 64, 2, 12
    net, Pd, cfg = _full_width_net(C.NCSNv2Deepest, "NCSNv2Deepest", n, 31, num_classes=levels, sigma_begin=30.0)
    sig = C.get_sigmas(cfg, mode="recons")
    A = C.SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 4, 1 / 8, seed=0)
    truth = phantom(1402, 1, 1, n, n)
    meas = A(truth.to(DEV)).repeat(1, B, 1, 1, 1)
    params = {"n_steps_each": 3, "step_lr": 9e-7, "denoise": True, "final_only": True}
    draw = lambda shape: torch.randn(*shape)
    sampler = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                              measurement=meas, linear_tfm=A, seg=None, device=torch.device(DEV))
    torch.manual_seed(78)
    got = sampler(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", noise_fn=draw)[0]
    torch.set_grad_enabled(True)
    maps, mask = A.sens_maps, A.random_under_fourier.mask
    score = lambda x, y: SN.score_forward("NCSNv2Deepest", Pd, x, y)
    prox = lambda z, y, a, l: M.l2_prox_sense_closed_form(z, y, maps, mask, a, l)
    torch.manual_seed(78)
    with torch.no_grad():
        ref = OALD.ald_sense_real_imag(score, meas.cpu(), sig.cpu(), 3, 9e-7, 1e6, lambda s_: M.sense_adjoint(s_, maps), prox)
    t = truth.abs()[0, 0]
    rng = float(t.max() - t.min())

    def metrics(x):
        mags = x.abs()[:, 0]
        mean_img = mags.mean(0)
        per = [(OALD.nrmse(m, t), OALD.ssim(m, t, data_range=rng)) for m in mags]
        return (OALD.nrmse(mean_img, t), OALD.ssim(mean_img, t, data_range=rng),
                sum(p_[0] for p_ in per) / len(per), sum(p_[1] for p_ in per) / len(per))
    mg, mr = metrics(got), metrics(ref)
    for a, b in zip(mg, mr):
        assert abs(a - b) < 1e-3, (mg, mr)
    assert rel_l2(got, ref) < 1e-3, rel_l2(got, ref)


@pytest.mark.parametrize("N,HW,Cc,offset", [(3, 32 * 32, 128, 0.0), (2, 64 * 64, 256, 50.0), (1, 17 * 9, 512, -3.0)])
def test_instnorm_plus_isolated(N, HW, Cc, offset):
    """InstanceNorm2dPlus (normalization.py:150-176) on its own: `ipdm_instnorm_stats` (pivoted and un-pivoted sums) and
    `ipdm_instnorm_apply_elu` (per-channel normalisation + the cross-channel m_hat term + gamma / beta + ELU) against the
    oracle; a large common offset makes the un-pivoted E[x^2] - E[x]^2 form lose digits, the pivoted one must not."""
    L = _lib()
    lib = L.lib()
    g = torch.Generator().manual_seed(N * 1000 + Cc)
    x = torch.randn(N, Cc, HW, 1, generator=g) * (0.5 + torch.rand(1, Cc, 1, 1, generator=g)) + torch.randn(1, Cc, 1, 1, generator=g) + offset
    Pn = {"n.alpha": 1 + 0.02 * torch.randn(Cc, generator=g), "n.gamma": 1 + 0.02 * torch.randn(Cc, generator=g),
          "n.beta": 0.1 * torch.randn(Cc, generator=g)}
    ref = torch.nn.functional.elu(SN.instance_norm_plus(Pn, "n", x.double()).float())          # (N, C, HW, 1)
    xd = x.reshape(N, Cc, HW).permute(0, 2, 1).contiguous().to(DEV)                             # NHWC
    a, gm, bt = (Pn[k].to(DEV) for k in ("n.alpha", "n.gamma", "n.beta"))
    for pivoted in (0, 1):
        stats = torch.zeros(N, Cc, 2, dtype=torch.float64, device=DEV)
        L.check(lib.ipdm_instnorm_stats(xd.data_ptr(), stats.data_ptr(), N, HW, Cc, pivoted, L.stream()), "stats")
        piv = xd[:, 0, :].double() if pivoted else torch.zeros(N, Cc, dtype=torch.float64, device=DEV)
        d = xd.double() - piv[:, None, :]
        assert rel_l2(stats[..., 0].cpu(), d.sum(1).cpu()) < 1e-5
        assert rel_l2(stats[..., 1].cpu(), (d * d).sum(1).cpu()) < 1e-5
        out = torch.empty(N, HW, Cc, dtype=torch.float16, device=DEV)
        L.check(lib.ipdm_instnorm_apply_elu(xd.data_ptr(), stats.data_ptr(), pivoted, a.data_ptr(), gm.data_ptr(), bt.data_ptr(),
                                            out.data_ptr(), N, HW, Cc, L.stream()), "apply")
        got = out.float().permute(0, 2, 1).reshape(N, Cc, HW, 1).cpu()
        tol = 1.5e-3 if (pivoted or offset == 0.0) else 2e-2      # f16 output rounding ~5e-4; un-pivoted sums at offset 50 lose ~3 digits
        assert rel_l2(got, ref) < tol, (pivoted, rel_l2(got, ref))
    # beta may be absent (the kernel accepts NULL)
    out = torch.empty(N, HW, Cc, dtype=torch.float16, device=DEV)
    L.check(lib.ipdm_instnorm_apply_elu(xd.data_ptr(), stats.data_ptr(), 1, a.data_ptr(), gm.data_ptr(), None, out.data_ptr(), N, HW, Cc,
                                        L.stream()), "apply")
    Pn.pop("n.beta")
    ref2 = torch.nn.functional.elu(SN.instance_norm_plus(Pn, "n", x.double()).float())
    assert rel_l2(out.float().permute(0, 2, 1).reshape(N, Cc, HW, 1).cpu(), ref2) < 1.5e-3


def test_chain_noise_is_keyed_by_global_chain_id():
    """chain i draws Philox(seed, chain i) -- the same chain whether it runs alone, at any slot of any batch, or
    on any rank of any world size.  Kernel level (bit-identical) for the fused step and the generic update, then through the
    public sampler: chain 5 alone == chain 5 inside the 14-chain batch of a single GPU == chain 5 on rank 5 of 8."""
    from inverseproblemwithdiffusionmodel_b200 import chains as CH
    L = _lib()
    lib = L.lib()
    n, nc = 128, 4
    A = C.SENSE("exp", nc, 40, 1 / 16, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 16, seed=0)
    mre, _ = A.device_maps(torch.device(DEV))
    plan = A.device_plan(torch.device(DEV), n)
    m8, frames = A.device_mask(torch.device(DEV))
    sc = L.AldScalars(0.1, 0.45, 0.02, 1.0)
    g = torch.Generator().manual_seed(5)
    per_chain = {i: (torch.randn(2, n, n, generator=g), torch.randn(2, n, n, generator=g), torch.randn(2, n, n, generator=g)) for i in range(16)}

    def run(ids, fused_plan=True):
        st, gr, bv = (torch.stack([per_chain[i][k] for i in ids], 1).contiguous().to(DEV) for k in range(3))
        idt = torch.tensor(ids, dtype=torch.int32, device=DEV)
        if fused_plan:
            L.check(lib.ipdm_ald_sense_step_plan(plan.handle, st.data_ptr(), gr.data_ptr(), None, bv.data_ptr(), mre.data_ptr(), None, nc, len(ids), n,
                                                 sc, None, None, L.rng(99, 7, idt), L.stream()), "step")
        else:
            L.check(lib.ipdm_ald_sense_step(st.data_ptr(), gr.data_ptr(), None, bv.data_ptr(), mre.data_ptr(), None, m8.data_ptr(), frames, nc, len(ids),
                                            n, n, sc, None, None, L.rng(99, 7, idt), L.stream()), "step")
        return st.cpu()
    for fused_plan in (True, False):
        alone = run([5], fused_plan)
        batch = run(list(range(14)), fused_plan)
        rank5 = run(CH.chain_partition(16, 8, 5), fused_plan)           # chains 5, 13
        assert torch.equal(alone[:, 0], batch[:, 5]) and torch.equal(alone[:, 0], rank5[:, 0])
        assert not torch.equal(batch[:, 5], batch[:, 6])
    # generic update (cfg 1 / sde corrector): per-sample streams
    xs = {i: torch.randn(1, 28, 28, generator=g) for i in range(16)}

    def run_l(ids):
        x = torch.stack([xs[i] for i in ids]).contiguous().to(DEV)
        gz = torch.zeros_like(x)
        idt = torch.tensor(ids, dtype=torch.int32, device=DEV)
        L.check(lib.ipdm_langevin_update(x.data_ptr(), gz.data_ptr(), None, None, x.numel(), L.AldScalars(0.0, 1.0, 0.0, 0.0), None, None, None, 0,
                                         L.rng(4, 2, idt, 28 * 28), L.stream()), "langevin")
        return x.cpu()
    assert torch.equal(run_l([5])[0], run_l(list(range(14)))[5]) and torch.equal(run_l([5])[0], run_l([13, 5])[1])
    # public sampler, captured-graph path with in-kernel noise
    n2 = 32
    cfg = C.make_config("ACDC", 8, n2, 10, 30.0, device=DEV)
    net, _ = C.build_net(C.NCSNv2Deepest, "NCSNv2Deepest_ngf8", 4, cfg, DEV)
    sig = C.get_sigmas(cfg, mode="recons")
    A2 = C.SENSE("exp", 4, 40, 1 / 8, (1, n2, n2), 0)
    A2.random_under_fourier.mask = C.keep_center_mask(n2, 4, 1 / 8, seed=0)
    y1 = A2(phantom(1403, 1, 1, n2, n2).to(DEV))
    params = {"n_steps_each": 2, "step_lr": 9e-7, "denoise": True, "final_only": True}

    def sample(ids):
        Bn = len(ids)
        s_ = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A2), 1.0, "linear", (Bn, 1, n2, n2), net, sig, params, cfg,
                                             measurement=y1.repeat(1, Bn, 1, 1, 1), linear_tfm=A2, seg=None, device=torch.device(DEV))
        out = s_(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", seed=2024, chain_ids=ids)[0]
        torch.set_grad_enabled(True)
        return out
    alone, batch, rank5 = sample([5]), sample(list(range(14))), sample(CH.chain_partition(105, 8, 5))
    assert rel_l2(alone[0], batch[5]) < 1e-6 and rel_l2(alone[0], rank5[0]) < 1e-6
    assert rel_l2(batch[5], batch[6]) > 1e-3
    # and without a seed two calls differ (fresh stream per call), with torch.manual_seed they repeat
    s_ = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A2), 1.0, "linear", (2, 1, n2, n2), net, sig, params, cfg,
                                         measurement=y1.repeat(1, 2, 1, 1, 1), linear_tfm=A2, seg=None, device=torch.device(DEV))
    kw = dict(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full")
    torch.manual_seed(1); a1 = s_(**kw)[0]; a2 = s_(**kw)[0]
    torch.manual_seed(1); a3 = s_(**kw)[0]
    torch.set_grad_enabled(True)
    assert rel_l2(a1, a2) > 1e-3 and torch.equal(a1, a3)


def test_fused_step_kernel_vs_oracle_256():
    """ipdm_ald_sense_step on a cfg-2 sized state (256x256, 4 coils, R=40) against the closed form."""
    L = _lib()
    n, B = 256, 3
    A = C.SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 64, seed=0)
    x = crandn(1, B, 1, n, n)
    g = crandn(2, B, 1, n, n)
    nz = crandn(3, B, 1, n, n)
    y = M.sense_forward(crandn(4, B, 1, n, n), A.sens_maps, A.random_under_fourier.mask)
    step, kappa = 0.37, 0.8
    ns = float(torch.sqrt(torch.tensor(step) * 2))
    z = x + step * g + ns * nz
    ref = M.l2_prox_sense_closed_form(z, y, A.sens_maps, A.random_under_fourier.mask, kappa * 4 * n / 0.05, 1.0)
    planar = lambda c: torch.stack([c.real.reshape(B, n, n), c.imag.reshape(B, n, n)]).contiguous().to(DEV)
    state, grad, noise = planar(x), planar(g), planar(nz)
    bvec = planar(M.sense_adjoint(y, A.sens_maps))
    mre, mim = A.device_maps(torch.device(DEV))
    m, frames = A.device_mask(torch.device(DEV))
    sc = L.AldScalars(step, ns, kappa, 0.0)
    L.check(L.lib().ipdm_ald_sense_step(state.data_ptr(), grad.data_ptr(), noise.data_ptr(), bvec.data_ptr(), mre.data_ptr(), None,
                                        m.data_ptr(), frames, 4, B, n, n, sc, None, None, None, L.stream()), "ald_sense_step")
    got = torch.complex(state[0], state[1]).cpu().reshape(B, 1, n, n)
    assert rel_l2(got, ref) < 1e-5
    assert rel_l2(got - z, ref - z) < 1e-4


def test_philox_noise_statistics_and_graph_replay():
    """In-kernel noise: N(0,1) moments, reproducible per (seed, step), different across steps; the
    captured-graph fast path equals the eager path step for step (same Philox keys)."""
    L = _lib()
    n = 1 << 20
    x = torch.zeros(n, device=DEV)
    g = torch.zeros(n, device=DEV)
    sc = L.AldScalars(0.5, 1.0, 0.0, 0.0)
    call = lambda buf, seed, k: L.check(L.lib().ipdm_langevin_update(buf.data_ptr(), g.data_ptr(), None, None, n, sc, None, None, None, 0, L.rng(seed, k), L.stream()), "langevin")
    call(x, 42, 0)
    assert abs(float(x.mean())) < 5e-3 and abs(float(x.var()) - 1) < 1e-2
    assert abs(float((x ** 4).mean()) - 3) < 0.1
    x2 = torch.zeros(n, device=DEV)
    call(x2, 42, 0)
    assert torch.equal(x, x2)
    x3 = torch.zeros(n, device=DEV)
    call(x3, 42, 1)
    assert abs(float((x * x3).mean())) < 5e-3
    # fast (graph) path vs eager path of the cfg-1 sampler with Philox noise
    cfg = C.make_config("MNIST", 8, 28, 10, 20.0, device=DEV)
    net, _ = C.build_net(C.NCSNv2, "NCSNv2_ngf8_28", 3, cfg, DEV)
    sig = C.get_sigmas(cfg)
    params = {"n_steps_each": 2, "step_lr": 6.2e-6, "denoise": True, "final_only": True}
    x0 = rrand(5, 2, 1, 28, 28)
    s = C.ALD.ALDUnconditionalSampler((2, 1, 28, 28), net, sig, params, cfg, device=torch.device(DEV))
    a = s(seed=9, x_init=x0, cuda_graph=True)[0]
    b = s(seed=9, x_init=x0, cuda_graph=False)[0]
    torch.set_grad_enabled(True)
    assert s.launches_per_step is not None and s.launches_per_step > 50
    assert rel_l2(a, b) < 1e-5


def test_in_kernel_noise_never_produces_nonfinite_values():
    """3e8 in-kernel normals (fused SENSE step and plain Langevin update): every one finite.  A uniform that can round
    to exactly 1.0 (24 random bits + 0.5 in fp32) gives radius 0 * inf = NaN about once per 1.7e7 draws -- invisible in
    short parity cases, fatal for a 6933-step chain (found by tools/run_posterior.py)."""
    L = _lib()
    n, B = 256, 14
    A = C.SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 64, seed=0)
    mre, mim = A.device_maps(torch.device(DEV))
    m, frames = A.device_mask(torch.device(DEV))
    state = torch.zeros(2, B, n, n, device=DEV)
    grad = torch.zeros_like(state)
    bvec = torch.zeros_like(state)
    sc = L.AldScalars(0.5, 1.0, 0.0, 0.0)
    steps = 160
    for k in range(steps):
        L.check(L.lib().ipdm_ald_sense_step(state.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), mre.data_ptr(), None,
                                            m.data_ptr(), frames, 4, B, n, n, sc, None, None, L.rng(77, k), L.stream()), "ald_sense_step")
    assert bool(torch.isfinite(state).all())
    v = float(state.var()) / steps            # kappa = 0: the prox is the identity, the state is a sum of unit normals
    assert abs(v - 1) < 1e-2, v
    x = torch.zeros(1 << 24, device=DEV)
    g = torch.zeros_like(x)
    for k in range(10):
        L.check(L.lib().ipdm_langevin_update(x.data_ptr(), g.data_ptr(), None, None, x.numel(), L.AldScalars(0.0, 1.0, 0.0, 0.0),
                                             None, None, None, 0, L.rng(5, k), L.stream()), "langevin")
    assert bool(torch.isfinite(x).all())
    assert abs(float(x.var()) / 10 - 1) < 1e-2


def test_sampler_uncond():
    C.case_sampler_uncond(DEV)


def test_sampler_sense():
    C.case_sampler_sense(DEV)


def test_sampler_cine():
    C.case_sampler_cine(DEV)


def test_sense_sampler_graph_path_matches_eager():
    n, B = 32, 3
    cfg = C.make_config("ACDC", 8, n, 10, 30.0, device=DEV)
    net, _ = C.build_net(C.NCSNv2Deepest, "NCSNv2Deepest_ngf8", 4, cfg, DEV)
    sig = C.get_sigmas(cfg, mode="recons")
    A = C.SENSE("exp", 4, 40, 1 / 8, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, 4, 1 / 8, seed=0)
    meas = A(phantom(1401, 1, 1, n, n).to(DEV)).repeat(1, B, 1, 1, 1)
    params = {"n_steps_each": 3, "step_lr": 9e-7, "denoise": True, "final_only": True}
    mk = lambda: C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                                 measurement=meas, linear_tfm=A, seg=None, device=torch.device(DEV))
    a = mk()(label=None, lamda=1., save_dir="/tmp", lr_scaled=1e6, seg_mode="full", seed=3, cuda_graph=True)[0]
    b = mk()(label=None, lamda=1., save_dir="/tmp", lr_scaled=1e6, seg_mode="full", seed=3, cuda_graph=False)[0]
    torch.set_grad_enabled(True)
    assert rel_l2(a, b) < 1e-5
    # chains do not interact: chain 0 of a 3-chain batch == the same chain run alone?  (Philox keys are
    # element-indexed within the batch, so only the first plane-offset-free chain is comparable)
    assert torch.isfinite(a.abs()).all()


def test_ncsn3d_shallow_temporal_prior():
    C.case_ncsn3d_shallow(DEV)


def test_sampler_cine_learned_temporal_prior():
    C.case_sampler_cine_diffusion(DEV)


def test_cine_random_shift_graph_path_matches_per_step_path():
    C.case_cine_diffusion_shift_graph(DEV)


def test_conv3d_via_slices_vs_torch():
    """One 3x3x3 dilated convolution = three slice-shifted launches of the 2-D tensor-core kernels, against
    torch.nn.functional.conv3d on the same f16-rounded operands (d = 1, 2: persistent halo kernel; d = 4: per-tap kernel)."""
    import torch.nn.functional as F
    L = _lib()
    P, X, T, Y, Cin, Cout = 3, 8, 24, 8, 128, 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(P, X, T, Y, Cin, generator=g).half()
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).half()          # (co, ci, kx, ky, kt)
    bias = torch.randn(Cout, generator=g)
    xd, bias_d = x.to(DEV), bias.to(DEV)
    for d in (1, 2, 4):
        acc = torch.full((P * X, T, Y, Cout), float("nan"), device=DEV)
        st = torch.zeros(P, Cout, 2, dtype=torch.float64, device=DEV)
        for order, (kx, first) in enumerate(((0, True), (2, False), (1, False))):
            w2 = w[:, :, kx].permute(0, 3, 2, 1).contiguous().reshape(Cout, 9, Cin).to(DEV)     # [co][(kt, ky)][ci]
            last = order == 2
            desc = L.ConvDesc(xd.data_ptr(), w2.data_ptr(), bias_d.data_ptr() if last else None, None if first else acc.data_ptr(),
                              acc.data_ptr(), None, st.data_ptr() if last else None, P * X, T, Y, Cin, Cout, 9, d, 0, X, (kx - 1) * d)
            L.check(L.lib().ipdm_conv_igemm(ctypes.byref(desc), L.stream()), "conv3d plane")
        ref = F.conv3d(x.float().permute(0, 4, 1, 3, 2), w.float(), bias, padding=d, dilation=d)      # (P, C, X, Y, T)
        ref = ref.permute(0, 2, 4, 3, 1).reshape(P * X, T, Y, Cout)
        assert rel_l2(acc.cpu(), ref) < 1e-5, d
        want = torch.stack([ref.reshape(P, -1, Cout).sum(1), (ref ** 2).reshape(P, -1, Cout).sum(1)], -1)
        assert rel_l2(st.float().cpu(), want) < 1e-5, d
        # the same convolution as ONE launch: 27-tap K loop, the three kx-planes fetched as three halo tiles
        w27 = w.permute(0, 2, 4, 3, 1).contiguous().reshape(Cout, 27, Cin).to(DEV)                   # [co][(kx, kt, ky)][ci]
        one = torch.full((P * X, T, Y, Cout), float("nan"), device=DEV)
        h16 = torch.full((P * X, T, Y, Cout), float("nan"), device=DEV, dtype=torch.float16)
        st2 = torch.zeros(P, Cout, 2, dtype=torch.float64, device=DEV)
        res = torch.randn(P * X, T, Y, Cout, generator=g).to(DEV)
        desc = L.ConvDesc(xd.data_ptr(), w27.data_ptr(), bias_d.data_ptr(), res.data_ptr(), one.data_ptr(), h16.data_ptr(), st2.data_ptr(),
                          P * X, T, Y, Cin, Cout, 27, d, 1, X, 0)
        L.check(L.lib().ipdm_conv_igemm(ctypes.byref(desc), L.stream()), "conv3d fused")
        assert rel_l2(one.cpu(), ref + res.cpu()) < 1e-5, d
        assert rel_l2(h16.float().cpu(), torch.nn.functional.elu(ref + res.cpu())) < 1e-3, d
        full = ref + res.cpu()
        want2 = torch.stack([full.reshape(P, -1, Cout).sum(1), (full ** 2).reshape(P, -1, Cout).sum(1)], -1)
        assert rel_l2(st2.float().cpu(), want2) < 1e-5, d


def test_metrics_and_result_files():
    C.case_metrics(DEV)


def test_posterior_stats_kernel():
    from inverseproblemwithdiffusionmodel_b200.chains import PosteriorStats
    x = crandn(5, 7, 1, 16, 16)
    st = PosteriorStats(256, torch.device(DEV))
    st.add(x[:3].to(DEV))
    st.add(x[3:].to(DEV))
    out = st.finalize((16, 16))
    ref = OALD.posterior_stats(x)
    for k in ("mag_mean", "mag_std", "phase_mean", "phase_std"):
        assert torch.allclose(out[k].cpu(), ref[k].reshape(16, 16), atol=2e-5), k
    assert out["n"] == 7


def test_full_chain_posterior_metrics():
    C.case_full_chain_metrics(DEV)


def test_map_baselines():
    C.case_map_baselines(DEV)


def test_deeper_langevin_seg():
    C.case_deeper_langevin_seg(DEV)
