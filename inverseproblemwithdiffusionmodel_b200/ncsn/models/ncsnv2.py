"""Mirror of `ncsn/models/ncsnv2.py` (+ the blocks of `layers.py` / `normalization.py` it uses):
`NCSNv2`, `NCSNv2Deeper` and `NCSNv2Deepest` RefineNet score networks with the reference's module tree, parameter
names, shapes and creation order (so `load_state_dict` of a reference checkpoint works and the
default initialisation consumes torch's RNG identically), but whose forward pass is a fixed
sequence of hand-written sm_100a kernels:

  * every 3x3 / 1x1 convolution with Cin % 64 == 0 and Cout % 128 == 0 (all of them at ngf = 128)
    runs as the tcgen05/TMEM implicit GEMM `ipdm_conv_igemm` on f16 NHWC operands with fp32
    accumulation; bias, residual add, ELU, the f16 copy for the next convolution, 2x2 mean-pooling
    and the InstanceNorm++ sums are epilogue work;
  * narrow nets (tests) use `ipdm_conv_direct`, which has the same epilogue contract;
  * begin_conv / end_conv (one channel on one side) are direct kernels with `2x-1` and `/sigma[y]` folded in;
  * InstanceNorm++ is one apply kernel (normalise + alpha*m_hat + gamma/beta + ELU -> f16 operand).

Convolution operands are f16.  The residual stream -- the un-activated tensors the blocks add
into -- is kept in f16 as well when every inner convolution runs on the tensor-core kernels (ngf a multiple of 128:
the real networks): a residual convolution then moves 8 instead of 12 bytes per output element, which takes the
dominant launch from under to over the ridge of the tensor / HBM roofline; all arithmetic (accumulators, residual adds,
InstanceNorm++ sums) stays fp32 and a value is rounded once, when stored.  Measured cost (oracle emulation and
kernels alike): the score's distance from the fp32 reference goes from 8e-4 to 1.2e-3 relative L2, an ALD step stays
below 1e-5 on x in the steady state.  Narrow test networks (CUDA-core convolutions) and IPDM_STREAM_F32=1 keep fp32.
"""
import ctypes
import os
from functools import partial

import torch
import torch.nn as nn

from . import get_sigmas
from ... import _lib
from ..._lib import ConvDesc, CONV_F16_ELU, CONV_F16_PRE_RES, CONV_RES_ELU, CONV_POOL2


# ------------------------------------------------------------------------------------------------
# parameter containers: same tree / names / creation order as the reference modules
# ------------------------------------------------------------------------------------------------
class InstanceNorm2dPlus(nn.Module):
    """parameters of normalization.py:150-176 (alpha, gamma ~ N(1, 0.02), beta = 0)"""

    def __init__(self, num_features, bias=True):
        super().__init__()
        self.num_features = num_features
        self.bias = bias
        self.instance_norm = nn.InstanceNorm2d(num_features, affine=False, track_running_stats=False)
        self.alpha = nn.Parameter(torch.zeros(num_features))
        self.gamma = nn.Parameter(torch.zeros(num_features))
        self.alpha.data.normal_(1, 0.02)
        self.gamma.data.normal_(1, 0.02)
        if bias:
            self.beta = nn.Parameter(torch.zeros(num_features))


def _conv3x3(cin, cout, bias=True, dilation=1):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=dilation, dilation=dilation, bias=bias)


class ConvMeanPool(nn.Module):
    """layers.py:291-313 (conv then mean of the four stride-2 phases)"""

    def __init__(self, input_dim, output_dim, kernel_size=3, biases=True):
        super().__init__()
        self.conv = nn.Conv2d(input_dim, output_dim, kernel_size, stride=1, padding=kernel_size // 2, bias=biases)


class ResidualBlock(nn.Module):
    """layers.py:401-456"""

    def __init__(self, input_dim, output_dim, resample=None, dilation=None):
        super().__init__()
        self.input_dim, self.output_dim, self.resample, self.dilation = input_dim, output_dim, resample, dilation
        if resample == 'down':
            if dilation is not None:
                self.conv1 = _conv3x3(input_dim, input_dim, dilation=dilation)
                self.normalize2 = InstanceNorm2dPlus(input_dim)
                self.conv2 = _conv3x3(input_dim, output_dim, dilation=dilation)
                shortcut = partial(_conv3x3, dilation=dilation)
            else:
                self.conv1 = _conv3x3(input_dim, input_dim)
                self.normalize2 = InstanceNorm2dPlus(input_dim)
                self.conv2 = ConvMeanPool(input_dim, output_dim, 3)
                shortcut = partial(ConvMeanPool, kernel_size=1)
        elif resample is None:
            if dilation is not None:
                shortcut = partial(_conv3x3, dilation=dilation)
                self.conv1 = _conv3x3(input_dim, output_dim, dilation=dilation)
                self.normalize2 = InstanceNorm2dPlus(output_dim)
                self.conv2 = _conv3x3(output_dim, output_dim, dilation=dilation)
            else:
                shortcut = lambda i, o: nn.Conv2d(i, o, kernel_size=1, stride=1, padding=0)
                self.conv1 = _conv3x3(input_dim, output_dim)
                self.normalize2 = InstanceNorm2dPlus(output_dim)
                self.conv2 = _conv3x3(output_dim, output_dim)
        else:
            raise Exception('invalid resample value')
        if output_dim != input_dim or resample is not None:
            self.shortcut = shortcut(input_dim, output_dim)
        self.normalize1 = InstanceNorm2dPlus(input_dim)


class RCUBlock(nn.Module):
    """layers.py:112-134 (bias-free convs named '{block}_{stage}_conv')"""

    def __init__(self, features, n_blocks, n_stages):
        super().__init__()
        for i in range(n_blocks):
            for j in range(n_stages):
                setattr(self, '{}_{}_conv'.format(i + 1, j + 1), _conv3x3(features, features, bias=False))
        self.n_blocks, self.n_stages = n_blocks, n_stages


class CRPBlock(nn.Module):
    """layers.py:62-83"""

    def __init__(self, features, n_stages):
        super().__init__()
        self.convs = nn.ModuleList([_conv3x3(features, features, bias=False) for _ in range(n_stages)])
        self.n_stages = n_stages


class MSFBlock(nn.Module):
    """layers.py:165-184"""

    def __init__(self, in_planes, features):
        super().__init__()
        self.convs = nn.ModuleList([_conv3x3(p, features, bias=True) for p in in_planes])
        self.features = features


class RefineBlock(nn.Module):
    """layers.py:214-249"""

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
class _Plan:
    """Buffers + packed weights of one network for one input shape (N, H, W) on one device."""

    def __init__(self, net, N, H, W, device):
        self.net, self.N, self.H, self.W, self.device = net, N, H, W, device
        self.bufs = {}
        self.w = {}
        self.version = None
        self._audited = None
        self.L = _lib.lib()
        inner = [m for name, m in net.named_modules() if isinstance(m, nn.Conv2d) and name not in ("begin_conv", "end_conv")]
        self.t16 = (len(inner) > 0 and all(m.in_channels % 64 == 0 and m.out_channels % 128 == 0 for m in inner)
                    and os.environ.get("IPDM_STREAM_F32") is None)      # (the 3-D network's plan keeps the fp32 stream)
        # Operand exponent shift (block floating point per tensor): f16 operands that do not come out of an InstanceNorm++
        # are stored as 2^-shift * value and the consuming convolution multiplies its accumulator by 2^shift -- exact, and
        # it moves the end of the f16 range from 6.5e4 to 6.5e4 * 2^shift.  0 unless the first-forward audit finds clipped
        # activations (`_auto_audit`), or IPDM_OPERAND_SHIFT says so; a shift > 0 also keeps the residual stream in fp32.
        self.shift = 0
        self.scale_of = {}       # data_ptr of an f16 operand tensor -> its scale (absent: 1)
        if os.environ.get("IPDM_OPERAND_SHIFT"):
            self.set_shift(int(os.environ["IPDM_OPERAND_SHIFT"]))

    def set_shift(self, shift):
        self.shift = int(shift)
        if self.shift > 0:
            self.t16 = False
        self.bufs.clear()
        self.scale_of.clear()

    @property
    def oscale(self):
        return 2.0 ** (-self.shift)

    # ---- storage ------------------------------------------------------------------------------
    def buf(self, name, shape, dtype):
        t = self.bufs.get(name)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.device)   # zeros, once: the range audit reads every f16 buffer whole
            self.bufs[name] = t
        return t

    def f32(self, name, N, H, W, C):
        """a tensor of the residual stream: fp32, or f16 on the tensor-core path (see the module docstring)"""
        return self.buf(name, (N, H, W, C), torch.float16 if self.t16 else torch.float32)

    def f16(self, name, N, H, W, C):
        return self.buf(name, (N, H, W, C), torch.float16)

    def stats(self, name, C):
        return self.buf(name, (self.N, C, 2), torch.float64)

    def pack(self):
        """(Re)build the kernel-side weight copies when any parameter changed."""
        net = self.net
        version = sum(p._version for p in net.parameters()) + sum(id(p) & 0xffff for p in net.parameters())
        if version == self.version:
            return
        s = _lib.stream()
        for name, mod in net.named_modules():
            if not isinstance(mod, nn.Conv2d):
                continue
            w = mod.weight.detach().to(self.device, torch.float32).contiguous()
            bias = None if mod.bias is None else mod.bias.detach().to(self.device, torch.float32).contiguous()
            cout, cin, kh, kw = w.shape
            if name == "begin_conv":
                if cin != 1:
                    raise _lib.IpdmError("begin_conv: only config.data.channels == 1 is implemented")
                self.w[name] = (w.reshape(cout, 9).contiguous(), bias)
            elif name == "end_conv":
                if cout != 1:
                    raise _lib.IpdmError("end_conv: only config.data.channels == 1 is implemented")
                self.w[name] = (w[0].permute(1, 2, 0).reshape(9, cin).contiguous(), bias)
            else:
                w16 = torch.empty((cout, kh * kw, cin), dtype=torch.float16, device=self.device)
                _lib.check(self.L.ipdm_pack_weights_f16(w.data_ptr(), w16.data_ptr(), cout, cin, kh * kw, s), "pack_weights")
                self.w[name] = (w16, bias)
                if name.endswith(".conv2.conv") and kh == 3 and 4 * cin <= 1024:
                    # ConvMeanPool (3x3 convolution, then 2x2 mean) = ONE 4x4 stride-2 convolution = a 3x3 convolution over the
                    # space-to-depth input [.., (y&1)*2 + (x&1), C] in which parity (pr, pc) meets only 2 x 2 of the 9 taps:
                    #   row offset dy of the s2d tensor, row parity pr  <-  original row taps oy:   (0,0): {0,-1}  (0,+1): {+1}
                    #                                                                                (1,0): {+1,0}  (1,-1): {-1}
                    R = torch.zeros(2, 3, 3, device=self.device)            # R[parity][dy + 1][o + 1]
                    R[0, 1, 1] = R[0, 1, 0] = R[0, 2, 2] = 1.0
                    R[1, 1, 2] = R[1, 1, 1] = R[1, 0, 0] = 1.0
                    w2 = 0.25 * torch.einsum("pyo,qxu,kcou->kpqcyx", R, R, w)   # [Cout][pr][pc][Cin][dy][dx]
                    w2 = w2.reshape(cout, 4 * cin, 3, 3).contiguous()
                    w16s = torch.empty((cout, 9, 4 * cin), dtype=torch.float16, device=self.device)
                    _lib.check(self.L.ipdm_pack_weights_f16(w2.data_ptr(), w16s.data_ptr(), cout, 4 * cin, 9, s), "pack_weights s2d")
                    mask = []
                    for kc in range(4 * cin // 64 if cin % 64 == 0 else 0):   # (narrow nets: no hint, the zero blocks are multiplied)
                        par = (kc * 64) // cin
                        pr, pc = par >> 1, par & 1
                        rows = (1, 2) if pr == 0 else (0, 1)                # dy + 1
                        cols = (1, 2) if pc == 0 else (0, 1)
                        mask.append(sum(1 << (3 * r + c) for r in rows for c in cols))
                    self.w[name + ".s2d"] = (w16s, bias, mask or None)
        for name, mod in net.named_modules():
            if isinstance(mod, InstanceNorm2dPlus):
                self.w[name] = tuple(None if t is None else t.detach().to(self.device, torch.float32).contiguous()
                                     for t in (mod.alpha, mod.gamma, mod.beta if mod.bias else None))
        self.sigmas = net.sigmas.detach().to(self.device, torch.float32).contiguous()
        self._sig_key = (id(net.sigmas), net.sigmas._version)
        self.version = version

    # ---- primitive launches ----------------------------------------------------------------------
    def conv(self, wname, x16, dims, residual=None, out32=None, out16=None, stats=None, flags=0, dilation=1):
        N, H, W, Cin, Cout = dims
        w16, bias = self.w[wname][:2]
        tap_mask = self.w[wname][2] if len(self.w[wname]) > 2 else None
        taps = w16.shape[1]
        if self.t16:     # residual / out32 are tensors of the 16-bit residual stream
            d = ConvDesc(_lib.ptr(x16), w16.data_ptr(), _lib.ptr(bias), None, None, _lib.ptr(out16),
                         _lib.ptr(stats), N, H, W, Cin, Cout, taps, dilation, flags, 0, 0, _lib.ptr(residual), _lib.ptr(out32))
        else:
            d = ConvDesc(_lib.ptr(x16), w16.data_ptr(), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(out32), _lib.ptr(out16),
                         _lib.ptr(stats), N, H, W, Cin, Cout, taps, dilation, flags)
        if tap_mask is not None:
            for i, m in enumerate(tap_mask):
                d.tap_mask[i] = m
        if self.shift:   # operand exponent shift: undo the input operand's scale on the accumulator, scale the operand output
            d.acc_scale = 1.0 / self.scale_of.get(x16.data_ptr(), 1.0)
            if out16 is not None:
                d.out_f16_scale = self.oscale
                self.scale_of[out16.data_ptr()] = self.oscale
        if Cin % 64 == 0 and Cout % 128 == 0:
            _lib.check(self.L.ipdm_conv_igemm(ctypes.byref(d), _lib.stream()), "conv_igemm " + wname)
        else:
            _lib.check(self.L.ipdm_conv_direct(ctypes.byref(d), _lib.stream()), "conv_direct " + wname)

    def norm_elu(self, nname, x32, stats, out16, N, HW, C):
        alpha, gamma, beta = self.w[nname]
        fn = self.L.ipdm_instnorm_apply_elu_f16in if self.t16 else self.L.ipdm_instnorm_apply_elu
        _lib.check(fn(x32.data_ptr(), stats.data_ptr(), 0, alpha.data_ptr(), gamma.data_ptr(),
                      _lib.ptr(beta), out16.data_ptr(), N, HW, C, _lib.stream()), "instnorm " + nname)

    def norm_elu_s2d(self, nname, x32, stats, out16, N, H, W, C):
        alpha, gamma, beta = self.w[nname]
        _lib.check(self.L.ipdm_instnorm_apply_elu_s2d(x32.data_ptr(), 1 if self.t16 else 0, stats.data_ptr(), 0, alpha.data_ptr(), gamma.data_ptr(),
                                                      _lib.ptr(beta), out16.data_ptr(), N, H, W, C, _lib.stream()), "instnorm s2d " + nname)

    def to_f16(self, x32, out16, elu):
        if self.shift:
            _lib.check(self.L.ipdm_act_to_f16_scaled(x32.data_ptr(), out16.data_ptr(), x32.numel(), 1 if elu else 0, self.oscale, _lib.stream()),
                       "act_to_f16_scaled")
            self.scale_of[out16.data_ptr()] = self.oscale
            return
        _lib.check(self.L.ipdm_act_to_f16(x32.data_ptr(), out16.data_ptr(), x32.numel(), 1 if elu else 0, _lib.stream()), "act_to_f16")

    def maxpool(self, x16, out16, N, H, W, C):
        _lib.check(self.L.ipdm_maxpool5_f16(x16.data_ptr(), out16.data_ptr(), N, H, W, C, _lib.stream()), "maxpool5")
        if self.shift:   # max commutes with a positive scale
            self.scale_of[out16.data_ptr()] = self.scale_of.get(x16.data_ptr(), 1.0)

    # ---- blocks -------------------------------------------------------------------------------------
    def residual_block(self, name, blk, x32, stats_x, H, W, elu16=None):
        """returns (out32, stats_out, H', W'); stats_* are un-pivoted per-(n,c) sums of the tensor.
        elu16 (optional f16 buffer) also receives f16(ELU(out)) for the decoder's skip connection."""
        N, Cin, Cout = self.N, blk.input_dim, blk.output_dim
        d = blk.dilation or 1
        pooled = blk.resample == 'down' and blk.dilation is None
        Cmid = blk.conv1.out_channels
        a1 = self.f16(name + ".a1", N, H, W, Cin)
        self.norm_elu(name + ".normalize1", x32, stats_x, a1, N, H * W, Cin)
        h32 = self.f32(name + ".h", N, H, W, Cmid)
        st_h = self.stats(name + ".st_h", Cmid)
        self.conv(name + ".conv1", a1, (N, H, W, Cin, Cmid), out32=h32, stats=st_h, dilation=d)
        # the strided forms of the pooled convolutions pay off when the POOLED image still fills the machine with pixel tiles
        # (tools/bench_pooled_conv.py: 28 x 256^2 0.83 -> 0.46 ms, 2 x 64^2 0.024 -> 0.040 ms): at least one wave of work items
        items = N * -(-(H // 2) // 32) * -(-(W // 2) // 8) * max(1, Cout // 128)
        big = (H // 2 >= 32 and items >= 148) or os.environ.get("IPDM_POOL_STRIDED_ALWAYS") is not None
        s2d = pooled and big and (name + ".conv2.conv.s2d") in self.w and os.environ.get("IPDM_POOL_AFTER_CONV") is None
        a2 = self.f16(name + ".a2", N, H, W, Cmid)
        if s2d:     # the same values in space-to-depth order: (N, H/2, W/2, 4*Cmid) in the same buffer
            self.norm_elu_s2d(name + ".normalize2", h32, st_h, a2, N, H, W, Cmid)
        else:
            self.norm_elu(name + ".normalize2", h32, st_h, a2, N, H * W, Cmid)
        Ho, Wo = (H // 2, W // 2) if pooled else (H, W)
        out32 = self.f32(name + ".out", N, Ho, Wo, Cout)
        st_o = self.stats(name + ".st_o", Cout)
        if hasattr(blk, "shortcut"):
            if self.t16:
                x16 = x32                                    # the 16-bit stream is the shortcut convolution's operand as it is
            else:
                x16 = self.f16(name + ".x16", N, H, W, Cin)
                self.to_f16(x32, x16, elu=False)
            sc = self.f32(name + ".sc", N, Ho, Wo, Cout)
            if pooled and big and Cin % 8 == 0 and self.w[name + ".shortcut.conv"][0].shape[1] == 1 and os.environ.get("IPDM_POOL_AFTER_SHORTCUT") is None:
                # mean-pool and the 1x1 shortcut convolution commute (the bias too): pool the operand, convolve a quarter of the pixels
                xp = self.f16(name + ".xp16", N, Ho, Wo, Cin)
                _lib.check(self.L.ipdm_meanpool2_f16(x16.data_ptr(), xp.data_ptr(), N, H, W, Cin, _lib.stream()), "meanpool2_f16")
                if self.shift:
                    self.scale_of[xp.data_ptr()] = self.scale_of.get(x16.data_ptr(), 1.0)
                self.conv(name + ".shortcut.conv", xp, (N, Ho, Wo, Cin, Cout), out32=sc)
            elif pooled:
                self.conv(name + ".shortcut.conv", x16, (N, H, W, Cin, Cout), out32=sc, flags=CONV_POOL2)
            else:
                k = blk.shortcut.kernel_size[0]
                self.conv(name + ".shortcut", x16, (N, H, W, Cin, Cout), out32=sc, dilation=d if k == 3 else 1)
            res = sc
        else:
            res = x32
        if s2d:
            self.conv(name + ".conv2.conv.s2d", a2, (N, Ho, Wo, 4 * Cmid, Cout), residual=res, out32=out32, out16=elu16, stats=st_o,
                      flags=CONV_F16_ELU)
        elif pooled:
            self.conv(name + ".conv2.conv", a2, (N, H, W, Cmid, Cout), residual=res, out32=out32, out16=elu16, stats=st_o,
                      flags=CONV_POOL2 | CONV_F16_ELU)
        else:
            self.conv(name + ".conv2", a2, (N, H, W, Cmid, Cout), residual=res, out32=out32, out16=elu16, stats=st_o,
                      flags=CONV_F16_ELU, dilation=d)
        return out32, st_o, Ho, Wo

    def rcu(self, name, blk, x32, e16, H, W, C, last_f16, last_stats=None):
        """x32 raw stream, e16 = f16(ELU(x32)).  last_f16: None | 'raw' | 'elu' -- the f16 copy the
        consumer of the block output wants; last_stats: InstanceNorm++ sums of the block output.
        Returns (out32, out16)."""
        N = self.N
        dims = (N, H, W, C, C)
        for i in range(blk.n_blocks):
            t16 = self.f16(f"{name}.t{i}", N, H, W, C)
            self.conv(f"{name}.{i + 1}_1_conv", e16, dims, out16=t16, flags=CONV_F16_ELU)
            final = i == blk.n_blocks - 1
            o32 = self.f32(f"{name}.o{i}", N, H, W, C)
            want = 'elu' if not final else last_f16
            if want == 'raw' and self.t16:
                o16 = None                                   # the 16-bit stream itself is the raw f16 copy
            else:
                o16 = self.f16(f"{name}.e{i}", N, H, W, C) if want else None
            self.conv(f"{name}.{i + 1}_2_conv", t16, dims, residual=x32, out32=o32, out16=o16,
                      stats=last_stats if final else None, flags=CONV_F16_ELU if want == 'elu' else 0)
            x32, e16 = o32, (o32 if want == 'raw' and self.t16 else o16)
        return x32, e16

    def crp(self, name, blk, h32, e16, H, W, C):
        """h32 raw input, e16 = f16(ELU(h32)).  Returns (x32, f16(ELU(x32)))."""
        N = self.N
        dims = (N, H, W, C, C)
        p16 = self.f16(name + ".p0", N, H, W, C)
        self.maxpool(e16, p16, N, H, W, C)
        x1 = self.f32(name + ".x1", N, H, W, C)
        path16 = self.f16(name + ".path1", N, H, W, C)
        self.conv(name + ".convs.0", p16, dims, residual=h32, out32=x1, out16=path16, flags=CONV_RES_ELU | CONV_F16_PRE_RES)
        p16b = self.f16(name + ".p1", N, H, W, C)
        self.maxpool(path16, p16b, N, H, W, C)
        x2 = self.f32(name + ".x2", N, H, W, C)
        e2 = self.f16(name + ".e2", N, H, W, C)
        self.conv(name + ".convs.1", p16b, dims, residual=x1, out32=x2, out16=e2, flags=CONV_F16_ELU)
        return x2, e2

    def refine(self, name, blk, inputs, H, W, final_stats=None):
        """inputs: list of (x32, e16, h, w, C).  Returns (out32, out16_elu)."""
        N, F = self.N, blk.features
        hs = []
        for i, (x32, e16, h, w, C) in enumerate(inputs):
            want = 'raw' if blk.n_blocks > 1 else 'elu'
            hs.append(self.rcu(f"{name}.adapt_convs.{i}", blk.adapt_convs[i], x32, e16, h, w, C, want) + (h, w, C))
        if blk.n_blocks > 1:
            sums = self.f32(name + ".sums", N, H, W, F)
            e16 = self.f16(name + ".sums_e", N, H, W, F)
            (a32, a16, ha, wa, Ca), (b32, b16, hb, wb, Cb) = hs
            assert (ha, wa) == (H, W)
            same = (hb, wb) == (H, W)
            self.conv(f"{name}.msf.convs.0", a16, (N, ha, wa, Ca, F), out32=sums)
            if same:
                self.conv(f"{name}.msf.convs.1", b16, (N, hb, wb, Cb, F), residual=sums, out32=sums, out16=e16, flags=CONV_F16_ELU)
            else:
                low = self.f32(name + ".low", N, hb, wb, F)
                self.conv(f"{name}.msf.convs.1", b16, (N, hb, wb, Cb, F), out32=low)
                fn = self.L.ipdm_bilinear_add_f16 if self.t16 else self.L.ipdm_bilinear_add
                if self.shift:   # the kernel's own f16(ELU) copy is unscaled: make the operand with the scaled cast instead
                    _lib.check(fn(low.data_ptr(), sums.data_ptr(), None, N, hb, wb, H, W, F, 1, _lib.stream()), "bilinear_add")
                    self.to_f16(sums, e16, elu=True)
                else:
                    _lib.check(fn(low.data_ptr(), sums.data_ptr(), e16.data_ptr(), N, hb, wb, H, W, F, 1, _lib.stream()), "bilinear_add")
            h32 = sums
        else:
            h32, e16 = hs[0][0], hs[0][1]
        x32, e16 = self.crp(name + ".crp", blk.crp, h32, e16, H, W, F)
        if final_stats is None:
            return self.rcu(name + ".output_convs", blk.output_convs, x32, e16, H, W, F, 'elu')
        # last refine block: the output feeds the final InstanceNorm++ -> needs its sums, no f16 copy
        return self.rcu(name + ".output_convs", blk.output_convs, x32, e16, H, W, F, None, last_stats=final_stats)

    def range_audit(self):
        """{buffer name: (max |x|, values at the end of the f16 range)} over every f16 buffer of the last forward, worst
        first.  f16 stores saturate at +-65504 (never inf / NaN), so an activation that left the range is silent in the
        output; this is how to find it (trained checkpoints are not available offline; default init peaks at 2.6e3)."""
        res = {}
        acc_m = torch.zeros(1, dtype=torch.float32, device=self.device)
        acc_n = torch.zeros(1, dtype=torch.int64, device=self.device)
        for name, t in self.bufs.items():
            if t.dtype != torch.float16:
                continue
            acc_m.zero_()
            acc_n.zero_()
            _lib.check(self.L.ipdm_f16_range_audit(t.data_ptr(), t.numel(), acc_m.data_ptr(), acc_n.data_ptr(), _lib.stream()), "range_audit")
            res[name] = (float(acc_m.item()), int(acc_n.item()))
        return dict(sorted(res.items(), key=lambda kv: -kv[1][0]))

    # ---- whole network ------------------------------------------------------------------------------
    def run(self, x, labels, out):
        """x f32 (N,1,H,W) contiguous, labels int64 (N,), out f32 (N,1,H,W) contiguous."""
        net, N, H, W = self.net, self.N, self.H, self.W
        self.pack()
        if (id(net.sigmas), net.sigmas._version) != self._sig_key:
            self.sigmas = net.sigmas.detach().to(self.device, torch.float32).contiguous()
            self._sig_key = (id(net.sigmas), net.sigmas._version)
        ngf = net.ngf
        s = _lib.stream()
        affine = 1 if (not net.logit_transform and not net.rescaled) else 0
        h32 = self.f32("begin", N, H, W, ngf)
        st = self.stats("begin.st", ngf)
        w0, b0 = self.w["begin_conv"]
        first = self.L.ipdm_conv_first_f16out if self.t16 else self.L.ipdm_conv_first
        _lib.check(first(x.data_ptr(), w0.data_ptr(), _lib.ptr(b0), h32.data_ptr(), st.data_ptr(), N, H, W, ngf, affine, s), "conv_first")
        feats = []
        ch, cw = H, W
        for stage in net.encoder_stages:
            blocks = getattr(net, stage)
            for i, blk in enumerate(blocks):
                e16 = None
                if i == len(blocks) - 1:
                    pooled = blk.resample == 'down' and blk.dilation is None
                    e16 = self.f16(stage + ".skip_e", N, ch // 2 if pooled else ch, cw // 2 if pooled else cw, blk.output_dim)
                h32, st, ch, cw = self.residual_block(f"{stage}.{i}", blk, h32, st, ch, cw, elu16=e16)
            feats.append((h32, e16, ch, cw, blocks[-1].output_dim))
        names = net.decoder_stages
        top = feats[-1]
        o32, o16 = self.refine(names[0], getattr(net, names[0]), [top], top[2], top[3])
        prev = (o32, o16, top[2], top[3], getattr(net, names[0]).features)
        fin = self.stats("final.st", ngf)
        for j, name in enumerate(names[1:], start=2):
            skip = feats[-j]
            blk = getattr(net, name)
            last = j == len(names)
            o32, o16 = self.refine(name, blk, [skip, prev], skip[2], skip[3], final_stats=fin if last else None)
            prev = (o32, o16, skip[2], skip[3], blk.features)
        a16 = self.f16("final.a", N, H, W, ngf)
        self.norm_elu("normalizer", prev[0], fin, a16, N, H * W, ngf)
        we, be = self.w["end_conv"]
        dots = self.buf("final.dots", (N, H, W, 9), torch.float32)
        _lib.check(self.L.ipdm_conv_last(a16.data_ptr(), we.data_ptr(), _lib.ptr(be), self.sigmas.data_ptr(), labels.data_ptr(),
                                         out.data_ptr(), dots.data_ptr(), N, H, W, ngf, s), "conv_last")
        self._auto_audit(x, labels, out)

    MAX_SHIFT = 18

    def _auto_audit(self, x, labels, out):
        """Once per set of weights (the first forward after they changed -- in the samplers that is the largest noise level,
        the largest activations): if an activation was clipped at the end of the f16 range, raise the operand exponent
        shift by 6 (x64 range; the residual stream goes to fp32), run the forward again and look again; give up with an
        error at 2^18.  Skipped inside a stream capture (the samplers run one eager step first);
        IPDM_ALLOW_F16_SATURATION=1 keeps the clipped result instead."""
        if self._audited == self.version or (torch.device(self.device).type == 'cuda' and torch.cuda.is_current_stream_capturing()):
            return
        self._audited = self.version
        if os.environ.get("IPDM_ALLOW_F16_SATURATION"):
            return
        worst = [(name, m, k) for name, (m, k) in self.range_audit().items() if k > 0]
        if not worst:
            return
        name, m, k = worst[0]
        what = (f"{sum(w[2] for w in worst)} activation values were clipped at +-65504 in the first forward with these weights "
                f"(worst buffer '{name}', {k} values, operand shift {self.shift})")
        if self.shift >= self.MAX_SHIFT:
            raise _lib.IpdmError("f16 range: " + what + f"; still clipped at the largest operand shift (2^{self.MAX_SHIFT})")
        import warnings
        warnings.warn("f16 range: " + what + f": re-running with operand exponent shift {self.shift + 6} and an fp32 residual stream "
                      "(DESIGN 2; slower: the straight-line epilogue paths need unscaled operands)")
        self.set_shift(self.shift + 6)
        self._audited = None
        self.run(x, labels, out)


class _ScoreNetBase(nn.Module):
    encoder_stages = ()
    decoder_stages = ()

    def _common(self, config):
        self.logit_transform = config.data.logit_transform
        self.rescaled = config.data.rescaled
        if config.model.normalization != 'InstanceNorm++' or config.model.nonlinearity.lower() != 'elu':
            raise _lib.IpdmError("only normalization 'InstanceNorm++' with nonlinearity 'elu' is implemented "
                                 "(what every config on the ALD path uses)")
        self.norm = InstanceNorm2dPlus
        self.ngf = config.model.ngf
        self.num_classes = config.model.num_classes
        self.register_buffer('sigmas', get_sigmas(config))
        self.config = config
        self._plans = {}

    def _plan(self, N, H, W, device):
        key = (N, H, W, device)
        plan = self._plans.get(key)
        if plan is None:
            plan = _Plan(self, N, H, W, device)
            self._plans[key] = plan
        return plan

    def range_audit(self):
        """Range audit of the f16 buffers of the most recent forward of every cached plan: {(N, H, W): {buffer: (max |x|,
        saturated values)}}; `saturated_total` > 0 means some activation was clipped to +-65504."""
        per_plan = {key[:3]: plan.range_audit() for key, plan in self._plans.items()}
        total = sum(v[1] for d in per_plan.values() for v in d.values())
        return {"plans": per_plan, "saturated_total": total,
                "max_abs": max((v[0] for d in per_plan.values() for v in d.values()), default=0.0)}

    def forward_into(self, x, labels, out):
        """No-allocation entry point for captured graphs: x, out f32 (N,1,H,W) contiguous CUDA buffers."""
        N, C, H, W = x.shape
        self._plan(N, H, W, x.device).run(x, labels, out)
        return out

    @torch.no_grad()
    def forward(self, x, y):
        """x: float32 (B, 1, H, W) (may be a non-contiguous view), y: int64 (B,) -> score (B, 1, H, W)"""
        _lib.require_cuda(x, y)
        if x.dim() != 4 or x.shape[1] != 1:
            raise _lib.IpdmError(f"score network expects (B, 1, H, W), got {tuple(x.shape)}")
        x = x.detach().to(torch.float32).contiguous()
        y = y.to(torch.int64).contiguous()
        out = torch.empty_like(x)
        return self.forward_into(x, y, out)


class NCSNv2(_ScoreNetBase):
    """ncsnv2.py:11-101: res1, res2 (down), res3 (dilation 2), res4 (dilation 4); four refine blocks."""
    encoder_stages = ("res1", "res2", "res3", "res4")
    decoder_stages = ("refine1", "refine2", "refine3", "refine4")

    def __init__(self, config):
        super().__init__()
        self._common(config)
        ngf = self.ngf
        self.begin_conv = nn.Conv2d(config.data.channels, ngf, 3, stride=1, padding=1)
        self.normalizer = self.norm(ngf)
        self.end_conv = nn.Conv2d(ngf, config.data.channels, 3, stride=1, padding=1)
        self.res1 = nn.ModuleList([ResidualBlock(ngf, ngf), ResidualBlock(ngf, ngf)])
        self.res2 = nn.ModuleList([ResidualBlock(ngf, 2 * ngf, resample='down'), ResidualBlock(2 * ngf, 2 * ngf)])
        self.res3 = nn.ModuleList([ResidualBlock(2 * ngf, 2 * ngf, resample='down', dilation=2),
                                   ResidualBlock(2 * ngf, 2 * ngf, dilation=2)])
        self.res4 = nn.ModuleList([ResidualBlock(2 * ngf, 2 * ngf, resample='down', dilation=4),
                                   ResidualBlock(2 * ngf, 2 * ngf, dilation=4)])
        self.refine1 = RefineBlock([2 * ngf], 2 * ngf, start=True)
        self.refine2 = RefineBlock([2 * ngf, 2 * ngf], 2 * ngf)
        self.refine3 = RefineBlock([2 * ngf, 2 * ngf], ngf)
        self.refine4 = RefineBlock([ngf, ngf], ngf, end=True)


class NCSNv2Deeper(_ScoreNetBase):
    """ncsnv2.py:104-195: two pooled stages, one plain-down stage pair, then dilation 2 / 4; five refine blocks."""
    encoder_stages = ("res1", "res2", "res3", "res4", "res5")
    decoder_stages = ("refine1", "refine2", "refine3", "refine4", "refine5")

    def __init__(self, config):
        super().__init__()
        self._common(config)
        ngf = self.ngf
        self.begin_conv = nn.Conv2d(config.data.channels, ngf, 3, stride=1, padding=1)
        self.normalizer = self.norm(ngf)
        self.end_conv = nn.Conv2d(ngf, config.data.channels, 3, stride=1, padding=1)
        self.res1 = nn.ModuleList([ResidualBlock(ngf, ngf), ResidualBlock(ngf, ngf)])
        self.res2 = nn.ModuleList([ResidualBlock(ngf, 2 * ngf, resample='down'), ResidualBlock(2 * ngf, 2 * ngf)])
        self.res3 = nn.ModuleList([ResidualBlock(2 * ngf, 2 * ngf, resample='down'), ResidualBlock(2 * ngf, 2 * ngf)])
        self.res4 = nn.ModuleList([ResidualBlock(2 * ngf, 4 * ngf, resample='down', dilation=2),
                                   ResidualBlock(4 * ngf, 4 * ngf, dilation=2)])
        self.res5 = nn.ModuleList([ResidualBlock(4 * ngf, 4 * ngf, resample='down', dilation=4),
                                   ResidualBlock(4 * ngf, 4 * ngf, dilation=4)])
        self.refine1 = RefineBlock([4 * ngf], 4 * ngf, start=True)
        self.refine2 = RefineBlock([4 * ngf, 4 * ngf], 2 * ngf)
        self.refine3 = RefineBlock([2 * ngf, 2 * ngf], 2 * ngf)
        self.refine4 = RefineBlock([2 * ngf, 2 * ngf], ngf)
        self.refine5 = RefineBlock([ngf, ngf], ngf, end=True)


class NCSNv2Deepest(_ScoreNetBase):
    """ncsnv2.py:198-299: three pooled stages then dilation 2 / 4; six refine blocks."""
    encoder_stages = ("res1", "res2", "res3", "res31", "res4", "res5")
    decoder_stages = ("refine1", "refine2", "refine31", "refine3", "refine4", "refine5")

    def __init__(self, config):
        super().__init__()
        self._common(config)
        ngf = self.ngf
        self.begin_conv = nn.Conv2d(config.data.channels, ngf, 3, stride=1, padding=1)
        self.normalizer = self.norm(ngf)
        self.end_conv = nn.Conv2d(ngf, config.data.channels, 3, stride=1, padding=1)
        self.res1 = nn.ModuleList([ResidualBlock(ngf, ngf), ResidualBlock(ngf, ngf)])
        self.res2 = nn.ModuleList([ResidualBlock(ngf, 2 * ngf, resample='down'), ResidualBlock(2 * ngf, 2 * ngf)])
        self.res3 = nn.ModuleList([ResidualBlock(2 * ngf, 2 * ngf, resample='down'), ResidualBlock(2 * ngf, 2 * ngf)])
        self.res31 = nn.ModuleList([ResidualBlock(2 * ngf, 2 * ngf, resample='down'), ResidualBlock(2 * ngf, 2 * ngf)])
        self.res4 = nn.ModuleList([ResidualBlock(2 * ngf, 4 * ngf, resample='down', dilation=2),
                                   ResidualBlock(4 * ngf, 4 * ngf, dilation=2)])
        self.res5 = nn.ModuleList([ResidualBlock(4 * ngf, 4 * ngf, resample='down', dilation=4),
                                   ResidualBlock(4 * ngf, 4 * ngf, dilation=4)])
        self.refine1 = RefineBlock([4 * ngf], 4 * ngf, start=True)
        self.refine2 = RefineBlock([4 * ngf, 4 * ngf], 2 * ngf)
        self.refine3 = RefineBlock([2 * ngf, 2 * ngf], 2 * ngf)
        self.refine31 = RefineBlock([2 * ngf, 2 * ngf], 2 * ngf)
        self.refine4 = RefineBlock([2 * ngf, 2 * ngf], ngf)
        self.refine5 = RefineBlock([ngf, ngf], ngf, end=True)
