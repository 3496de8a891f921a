"""Mirror of `ncsn/models/ncsn3d.py:123-224` (`NCSN3DShallow`, the learned temporal prior of the 2D+time sampler) with
the blocks of `layers3d.py` / `normalization3d.py` it uses: same module tree, parameter names, shapes and creation
order as the reference (so `load_state_dict` of a reference checkpoint works and default initialisation consumes
torch's RNG identically); the forward pass is a fixed sequence of sm_100a kernels.

Volumes are `(B, 1, kx, ky, T)` patches (8 x 8 x 24 for CINE127).  On the device they live as `[P][X][T][Y][C]`:
every X-slice is an NHWC "image" of H = T rows and W = Y columns, so

  * a 3x3x3 (dilated) convolution = ONE launch of the 2-D tcgen05 implicit GEMM with a 27-tap K loop: per 64 input
    channels the three kx-planes are three halo tiles fetched from slices x + (kx-1)*dilation of the volume
    (`ipdm_conv_desc.taps = 27, slices = X`; the 5-D TMA box zero-pads across slices) and all 27 taps accumulate in
    the same TMEM tile, so bias / residual / ELU / f16 copy / InstanceNorm++ sums are the ordinary epilogue;
  * InstanceNorm3dPlus is the 2-D apply kernel over X*T*Y positions; MaxPool3d(5) = the 2-D 5x5 pool + a slice-axis pool;
  * `conv_temporal_down` (Conv3d (1,1,4)/s(1,1,2)) and `conv_temporal_up` (ConvTranspose3d) are a T-gather that lays
    their taps side by side + ONE 1x1 implicit GEMM (the transposed one computes its two output phases as 2*Cout
    channels, interleaved back onto the time axis afterwards);
  * trilinear interpolation inside the refine blocks is the identity here (all inputs already have the output size).
"""
import ctypes
from functools import partial

import torch
import torch.nn as nn

from . import get_sigmas
from ... import _lib
from ..._lib import ConvDesc, CONV_F16_ELU
from .ncsnv2 import _Plan


# ------------------------------------------------------------------------------------------------
# parameter containers (reference tree / names / creation order)
# ------------------------------------------------------------------------------------------------
class InstanceNorm3dPlus(nn.Module):
    """parameters of normalization3d.py:151-185"""

    def __init__(self, num_features, bias=True):
        super().__init__()
        self.num_features = num_features
        self.bias = bias
        self.instance_norm = nn.InstanceNorm3d(num_features, affine=False, track_running_stats=False)
        self.alpha = nn.Parameter(torch.zeros(num_features))
        self.gamma = nn.Parameter(torch.zeros(num_features))
        self.alpha.data.normal_(1, 0.02)
        self.gamma.data.normal_(1, 0.02)
        if bias:
            self.beta = nn.Parameter(torch.zeros(num_features))


def _conv3(cin, cout, bias=True, dilation=1):
    return nn.Conv3d(cin, cout, kernel_size=3, stride=1, padding=dilation, dilation=dilation, bias=bias)


class ResidualBlock(nn.Module):
    """layers3d.py:423-476 (the un-pooled variants NCSN3DShallow instantiates)"""

    def __init__(self, input_dim, output_dim, resample=None, dilation=None):
        super().__init__()
        self.input_dim, self.output_dim, self.resample, self.dilation = input_dim, output_dim, resample, dilation
        if resample == 'down':
            if dilation is None:
                raise _lib.IpdmError("3-D ConvMeanPool residual blocks are not used by NCSN3DShallow and not implemented")
            self.conv1 = _conv3(input_dim, input_dim, dilation=dilation)
            self.normalize2 = InstanceNorm3dPlus(input_dim)
            self.conv2 = _conv3(input_dim, output_dim, dilation=dilation)
            shortcut = partial(_conv3, dilation=dilation)
        elif resample is None:
            if dilation is not None:
                shortcut = partial(_conv3, dilation=dilation)
                self.conv1 = _conv3(input_dim, output_dim, dilation=dilation)
                self.normalize2 = InstanceNorm3dPlus(output_dim)
                self.conv2 = _conv3(output_dim, output_dim, dilation=dilation)
            else:
                shortcut = lambda i, o: nn.Conv3d(i, o, kernel_size=1, stride=1, padding=0)
                self.conv1 = _conv3(input_dim, output_dim)
                self.normalize2 = InstanceNorm3dPlus(output_dim)
                self.conv2 = _conv3(output_dim, output_dim)
        else:
            raise Exception('invalid resample value')
        if output_dim != input_dim or resample is not None:
            self.shortcut = shortcut(input_dim, output_dim)
        self.normalize1 = InstanceNorm3dPlus(input_dim)


class RCUBlock(nn.Module):
    """layers3d.py:113-135"""

    def __init__(self, features, n_blocks, n_stages):
        super().__init__()
        for i in range(n_blocks):
            for j in range(n_stages):
                setattr(self, '{}_{}_conv'.format(i + 1, j + 1), _conv3(features, features, bias=False))
        self.n_blocks, self.n_stages = n_blocks, n_stages


class CRPBlock(nn.Module):
    """layers3d.py:63-84 (MaxPool3d(5, 1, 2))"""

    def __init__(self, features, n_stages):
        super().__init__()
        self.convs = nn.ModuleList([_conv3(features, features, bias=False) for _ in range(n_stages)])
        self.n_stages = n_stages


class MSFBlock(nn.Module):
    """layers3d.py:166-187"""

    def __init__(self, in_planes, features):
        super().__init__()
        self.convs = nn.ModuleList([_conv3(p, features, bias=True) for p in in_planes])
        self.features = features


class RefineBlock(nn.Module):
    """layers3d.py:219-255"""

    def __init__(self, in_planes, features, start=False, end=False):
        super().__init__()
        self.n_blocks = len(in_planes)
        self.in_planes, self.features, self.end = list(in_planes), features, end
        self.adapt_convs = nn.ModuleList([RCUBlock(p, 2, 2) for p in in_planes])
        self.output_convs = RCUBlock(features, 3 if end else 1, 2)
        if not start:
            self.msf = MSFBlock(in_planes, features)
        self.crp = CRPBlock(features, 2)


# ------------------------------------------------------------------------------------------------
# the kernel sequence
# ------------------------------------------------------------------------------------------------
class _Plan3D(_Plan):
    """Buffers + packed weights for one input shape: P volumes of X slices, each a (T x Y) image.  Reuses the block
    sequences of the 2-D plan (`residual_block`, `rcu`, `crp`, `refine`) with 3-D primitives underneath; in the
    inherited code `N` is the number of SLICES (P*X), `H` = T and `W` = Y."""

    def __init__(self, net, P, X, T, Y, device):
        super().__init__(net, P * X, T, Y, device)
        self.P, self.X = P, X

    def stats(self, name, C):
        return self.buf(name, (self.P, C, 2), torch.float64)

    def pack(self):
        net = self.net
        version = sum(p._version for p in net.parameters()) + sum(id(p) & 0xffff for p in net.parameters())
        if version == self.version:
            return
        s = _lib.stream()
        f32 = lambda t: None if t is None else t.detach().to(self.device, torch.float32).contiguous()

        def pack2d(w_oihw):
            cout, cin, kh, kw = w_oihw.shape
            w16 = torch.empty((cout, kh * kw, cin), dtype=torch.float16, device=self.device)
            w_oihw = w_oihw.contiguous()
            _lib.check(self.L.ipdm_pack_weights_f16(w_oihw.data_ptr(), w16.data_ptr(), cout, cin, kh * kw, s), "pack_weights")
            return w16

        for name, mod in net.named_modules():
            if isinstance(mod, nn.ConvTranspose3d):
                # (Cin, Cout, 1, 1, 4), stride 2, padding 1: out[2m] = W1^T x[m] + W3^T x[m-1], out[2m+1] = W2^T x[m] + W0^T x[m+1]
                w = f32(mod.weight)[:, :, 0, 0, :]                       # (Cin, Cout, 4)
                cin, cout, _ = w.shape
                z = torch.zeros(cout, cin, device=self.device)
                even = torch.cat([w[:, :, 3].t(), w[:, :, 1].t(), z], 1)   # gathered operand = [x[m-1], x[m], x[m+1]]
                odd = torch.cat([z, w[:, :, 2].t(), w[:, :, 0].t()], 1)
                w2 = torch.cat([even, odd], 0).reshape(2 * cout, 3 * cin, 1, 1)
                b = f32(mod.bias)
                self.w[name] = (pack2d(w2), None if b is None else torch.cat([b, b]).contiguous())
            elif isinstance(mod, nn.Conv3d):
                w, bias = f32(mod.weight), f32(mod.bias)
                cout, cin, kx, ky, kt = w.shape
                if name == "begin_conv":
                    if cin != 1:
                        raise _lib.IpdmError("begin_conv: only config.data.channels_3d == 1 is implemented")
                    self.w[name] = (w[:, 0].permute(0, 1, 3, 2).reshape(cout, 27).contiguous(), bias)        # taps (kx, kt, ky)
                elif name == "end_conv":
                    if cout != 1:
                        raise _lib.IpdmError("end_conv: only config.data.channels_3d == 1 is implemented")
                    self.w[name] = (w[0].permute(1, 3, 2, 0).reshape(27, cin).contiguous(), bias)            # [(kx, kt, ky)][C]
                elif (kx, ky, kt) == (3, 3, 3):
                    # [Cout][27][Cin], taps ordered (kx, kh = kt, kw = ky): the (T, Y) image of every kx-plane
                    w27 = w.permute(0, 1, 2, 4, 3).reshape(cout, cin, 27, 1)
                    self.w[name] = (pack2d(w27), bias)
                elif (kx, ky, kt) == (1, 1, 1):
                    self.w[name] = (pack2d(w[:, :, 0]), bias)
                elif (kx, ky) == (1, 1):
                    # strided temporal convolution: taps side by side, w2[co][k*Cin + ci] = w[co][ci][k]
                    self.w[name] = (pack2d(w[:, :, 0, 0, :].permute(0, 2, 1).reshape(cout, kt * cin, 1, 1)), bias)
                else:
                    raise _lib.IpdmError(f"{name}: unsupported Conv3d kernel {(kx, ky, kt)}")
        for name, mod in net.named_modules():
            if isinstance(mod, InstanceNorm3dPlus):
                self.w[name] = tuple(f32(t) for t in (mod.alpha, mod.gamma, mod.beta if mod.bias else None))
        self.sigmas = net.sigmas.detach().to(self.device, torch.float32).contiguous()
        self._sig_key = (id(net.sigmas), net.sigmas._version)
        self.version = version

    # ---- primitive launches ----------------------------------------------------------------------
    def _launch(self, w16, bias, x16, dims, residual, out32, out16, stats, flags, dilation, shift, what):
        N, H, W, Cin, Cout = dims
        if Cin % 64 or Cout % 128:
            raise _lib.IpdmError(f"{what}: the 3-D network needs Cin % 64 == 0 and Cout % 128 == 0 (ngf = 128), got {Cin} -> {Cout}")
        d = ConvDesc(_lib.ptr(x16), w16.data_ptr(), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(out32), _lib.ptr(out16),
                     _lib.ptr(stats), N, H, W, Cin, Cout, w16.shape[1], dilation, flags, self.X, shift)
        _lib.check(self.L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()), what)

    def conv(self, wname, x16, dims, residual=None, out32=None, out16=None, stats=None, flags=0, dilation=1):
        w16, bias = self.w[wname]
        taps = w16.shape[1]           # 27: 3x3x3 over the volume; 1: shortcut of a plain block / gathered temporal convolutions
        self._launch(w16, bias, x16, dims, residual, out32, out16, stats, flags, dilation if taps == 27 else 1, 0,
                     ("conv3d " if taps == 27 else "conv1x1 ") + wname)

    def norm_elu(self, nname, x32, stats, out16, N, HW, C):
        super().norm_elu(nname, x32, stats, out16, self.P, (N // self.P) * HW, C)

    def maxpool(self, x16, out16, N, H, W, C):
        tmp = self.f16("maxpool3d.tmp%d_%d" % (H, C), N, H, W, C)
        _lib.check(self.L.ipdm_maxpool5_f16(x16.data_ptr(), tmp.data_ptr(), N, H, W, C, _lib.stream()), "maxpool5")
        _lib.check(self.L.ipdm_maxpool5_slices_f16(tmp.data_ptr(), out16.data_ptr(), self.P, self.X, H * W * C, _lib.stream()),
                   "maxpool5_slices")

    # ---- whole network (ncsn3d.py:186-224) ------------------------------------------------------------
    def run(self, x, labels, out):
        """x, out f32 [P][X][T][Y] contiguous, labels int64 (P,)."""
        net, P, X, T, Y, N = self.net, self.P, self.X, self.H, self.W, self.N
        self.pack()
        if (id(net.sigmas), net.sigmas._version) != self._sig_key:
            self.sigmas = net.sigmas.detach().to(self.device, torch.float32).contiguous()
            self._sig_key = (id(net.sigmas), net.sigmas._version)
        ngf, s, L = net.ngf, _lib.stream(), self.L
        affine = 1 if (not net.logit_transform and not net.rescaled) else 0
        h32 = self.f32("begin", N, T, Y, ngf)
        st = self.stats("begin.st", ngf)
        w0, b0 = self.w["begin_conv"]
        _lib.check(L.ipdm_conv3d_first(x.data_ptr(), w0.data_ptr(), _lib.ptr(b0), h32.data_ptr(), P, X, T, Y, ngf, affine, s), "conv3d_first")
        _lib.check(L.ipdm_instnorm_stats(h32.data_ptr(), st.data_ptr(), P, X * T * Y, ngf, 0, s), "instnorm_stats")
        # encoder: res1 (T), res3 (dilation 2)
        e1 = self.f16("res1.skip_e", N, T, Y, ngf)
        h32, st, _, _ = self.residual_block("res1.0", net.res1[0], h32, st, T, Y)
        l1, st, _, _ = self.residual_block("res1.1", net.res1[1], h32, st, T, Y, elu16=e1)
        h32, st, _, _ = self.residual_block("res3.0", net.res3[0], l1, st, T, Y)
        l2, st, _, _ = self.residual_block("res3.1", net.res3[1], h32, st, T, Y)
        # conv_temporal_down: T -> T/2
        T2, C2 = T // 2, 2 * ngf
        l2_16 = self.f16("l2.raw16", N, T, Y, C2)
        self.to_f16(l2, l2_16, elu=False)
        g = self.f16("tdown.gather", N, T2, Y, 4 * C2)
        _lib.check(L.ipdm_gather_t_f16(l2_16.data_ptr(), g.data_ptr(), N, T, T2, Y, C2, 2, -1, 4, s), "gather_t")
        l3 = self.f32("l3", N, T2, Y, C2)
        e3 = self.f16("l3.e", N, T2, Y, C2)
        st3 = self.stats("l3.st", C2)
        self.conv("conv_temporal_down", g, (N, T2, Y, 4 * C2, C2), out32=l3, out16=e3, stats=st3, flags=CONV_F16_ELU)
        # res4 (dilation 4) at T/2
        e4 = self.f16("res4.skip_e", N, T2, Y, C2)
        h32, st, _, _ = self.residual_block("res4.0", net.res4[0], l3, st3, T2, Y)
        l4, st, _, _ = self.residual_block("res4.1", net.res4[1], h32, st, T2, Y, elu16=e4)
        # decoder
        r1, r1e = self.refine("refine1", net.refine1, [(l4, e4, T2, Y, C2)], T2, Y)
        r2, r2e = self.refine("refine2", net.refine2, [(l3, e3, T2, Y, C2), (r1, r1e, T2, Y, C2)], T2, Y)
        # conv_temporal_up: T/2 -> T, 2*ngf -> ngf
        r2_16 = self.f16("r2.raw16", N, T2, Y, C2)
        self.to_f16(r2, r2_16, elu=False)
        g2 = self.f16("tup.gather", N, T2, Y, 3 * C2)
        _lib.check(L.ipdm_gather_t_f16(r2_16.data_ptr(), g2.data_ptr(), N, T2, T2, Y, C2, 1, -1, 3, s), "gather_t")
        ph = self.f32("tup.phases", N, T2, Y, 2 * ngf)
        self.conv("conv_temporal_up", g2, (N, T2, Y, 3 * C2, 2 * ngf), out32=ph)
        r3 = self.f32("r3", N, T, Y, ngf)
        r3e = self.f16("r3.e", N, T, Y, ngf)
        _lib.check(L.ipdm_interleave_t(ph.data_ptr(), r3.data_ptr(), r3e.data_ptr(), N, T2, Y, ngf, s), "interleave_t")
        fin = self.stats("final.st", ngf)
        o32, _ = self.refine("refine3", net.refine3, [(l1, e1, T, Y, ngf), (r3, r3e, T, Y, ngf)], T, Y, final_stats=fin)
        a16 = self.f16("final.a", N, T, Y, ngf)
        self.norm_elu("normalizer", o32, fin, a16, N, T * Y, ngf)
        we, be = self.w["end_conv"]
        _lib.check(L.ipdm_conv3d_last(a16.data_ptr(), we.data_ptr(), _lib.ptr(be), self.sigmas.data_ptr(), labels.data_ptr(),
                                      out.data_ptr(), P, X, T, Y, ngf, s), "conv3d_last")


class NCSN3DShallow(nn.Module):
    """ncsn3d.py:123-224: res1, res3 (dilation 2), temporal stride-2 convolution, res4 (dilation 4), three refine blocks,
    transposed temporal convolution.  Input `(B, 1, kx, ky, T)` or the flattened `(B, kx*ky, T)`; `y` int64 `(B,)`."""

    def __init__(self, config):
        super().__init__()
        self.logit_transform = config.data.logit_transform
        self.rescaled = config.data.rescaled
        if config.model.normalization != 'InstanceNorm++' or config.model.nonlinearity.lower() != 'elu':
            raise _lib.IpdmError("only normalization 'InstanceNorm++' with nonlinearity 'elu' is implemented")
        self.norm = InstanceNorm3dPlus
        self.ngf = ngf = config.model.ngf
        self.num_classes = config.model.num_classes
        self.register_buffer('sigmas', get_sigmas(config))
        self.config = config
        self.begin_conv = nn.Conv3d(config.data.channels_3d, ngf, 3, stride=1, padding=1)
        self.normalizer = self.norm(ngf)
        self.end_conv = nn.Conv3d(ngf, config.data.channels_3d, 3, stride=1, padding=1)
        self.res1 = nn.ModuleList([ResidualBlock(ngf, ngf), ResidualBlock(ngf, ngf)])
        self.res3 = nn.ModuleList([ResidualBlock(ngf, 2 * ngf, resample='down', dilation=2),
                                   ResidualBlock(2 * ngf, 2 * ngf, dilation=2)])
        self.res4 = nn.ModuleList([ResidualBlock(2 * ngf, 2 * ngf, resample='down', dilation=4),
                                   ResidualBlock(2 * ngf, 2 * ngf, dilation=4)])
        self.refine1 = RefineBlock([2 * ngf], 2 * ngf, start=True)
        self.refine2 = RefineBlock([2 * ngf, 2 * ngf], 2 * ngf)
        self.refine3 = RefineBlock([ngf, ngf], ngf)
        self.conv_temporal_down = nn.Conv3d(2 * ngf, 2 * ngf, kernel_size=(1, 1, 4), stride=(1, 1, 2), padding=(0, 0, 1))
        self.conv_temporal_up = nn.ConvTranspose3d(2 * ngf, ngf, kernel_size=(1, 1, 4), stride=(1, 1, 2), padding=(0, 0, 1))
        self._plans = {}

    def _plan(self, P, X, T, Y, device):
        key = (P, X, T, Y, device)
        plan = self._plans.get(key)
        if plan is None:
            plan = _Plan3D(self, P, X, T, Y, device)
            self._plans[key] = plan
        return plan

    def forward_into(self, x_pxty, labels, out_pxty):
        """No-allocation entry point: x, out f32 [P][X][T][Y] contiguous CUDA buffers (the device layout)."""
        P, X, T, Y = x_pxty.shape
        self._plan(P, X, T, Y, x_pxty.device).run(x_pxty, labels, out_pxty)
        return out_pxty

    @torch.no_grad()
    def forward(self, x, y):
        _lib.require_cuda(x, y)
        flat = x.dim() == 3
        if flat:
            k = int(round(x.shape[1] ** 0.5))
            if k * k != x.shape[1]:
                raise _lib.IpdmError(f"(B, kx*ky, T) input needs a square patch, got {tuple(x.shape)}")
            x = x.reshape(x.shape[0], 1, k, k, x.shape[2])
        if x.dim() != 5 or x.shape[1] != 1:
            raise _lib.IpdmError(f"NCSN3DShallow expects (B, 1, kx, ky, T) or (B, kx*ky, T), got {tuple(x.shape)}")
        B, _, X, Y, T = x.shape
        if T % 2:
            raise _lib.IpdmError("the temporal stride-2 convolution needs an even T")
        xin = x.detach().to(torch.float32)[:, 0].permute(0, 1, 3, 2).contiguous()          # [P][X][T][Y]
        out = torch.empty_like(xin)
        self.forward_into(xin, y.to(torch.int64).contiguous(), out)
        res = out.permute(0, 1, 3, 2).unsqueeze(1).contiguous()                            # (B, 1, kx, ky, T)
        return res.reshape(B, X * Y, T) if flat else res
