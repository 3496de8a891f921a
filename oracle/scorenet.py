"""Oracle (test infrastructure): functional NCSNv2 / NCSNv2Deepest forward on a reference-layout state dict.

Restates `ncsn/models/ncsnv2.py`, `layers.py` and `normalization.py` of the reference as pure
functions over a `dict[str, Tensor]` whose keys are the reference's state-dict names
(`res1.0.conv1.weight`, `refine5.msf.convs.0.bias`, ...), using torch CPU ops.
"""
import torch
import torch.nn.functional as F


# None = exact fp32 reference arithmetic.  torch.float16 = additionally round the operands (input and
# weights) of every convolution with more than one channel on both sides to f16, fp32 accumulate --
# the arithmetic of the tensor-core path; used by tests to separate "logic
# differs" (tight tolerance against this mode) from "operand precision differs" (looser, vs fp32).
OPERAND_ROUND = None
# None = fp32 residual stream.  torch.float16 = additionally round every tensor of the residual stream (block outputs,
# pre-norm convolution results, shortcut results, MSF sums) when it is "stored", as the tensor-core path does when it keeps
# that stream in 16 bits; arithmetic stays fp32.
STREAM_ROUND = None


def _st(x):
    return x if STREAM_ROUND is None else x.to(STREAM_ROUND).float()


def _conv(P, key, x, dilation=1, kernel=3):
    """3x3 (pad = dilation) or 1x1 convolution, bias iff present in the state dict.
    Reference: conv3x3 / dilated_conv3x3 / conv1x1, layers.py:28-60."""
    w = P[key + ".weight"]
    b = P.get(key + ".bias")
    pad = dilation if kernel == 3 else 0
    if OPERAND_ROUND is not None and w.shape[0] > 1 and w.shape[1] > 1:
        x = x.to(OPERAND_ROUND).float()
        w = w.to(OPERAND_ROUND).float()
    return F.conv2d(x, w, b, stride=1, padding=pad, dilation=dilation)


def instance_norm_plus(P, key, x, eps=1e-5):
    """InstanceNorm++: gamma*(IN(x) + alpha*m_hat) + beta, m_hat = channel-standardised spatial
    means (unbiased variance over C). Reference: InstanceNorm2dPlus.forward, normalization.py:163-176."""
    mu = x.mean(dim=(2, 3))
    m_hat = (mu - mu.mean(dim=-1, keepdim=True)) / torch.sqrt(mu.var(dim=-1, keepdim=True) + eps)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    h = (x - mu[..., None, None]) / torch.sqrt(var + eps)
    h = h + (m_hat * P[key + ".alpha"])[..., None, None]
    out = P[key + ".gamma"].view(1, -1, 1, 1) * h
    if key + ".beta" in P:
        out = out + P[key + ".beta"].view(1, -1, 1, 1)
    return out


def mean_pool2(x):
    """Average of the four stride-2 phases. Reference: ConvMeanPool.forward, layers.py:309-313."""
    return (x[:, :, ::2, ::2] + x[:, :, 1::2, ::2] + x[:, :, ::2, 1::2] + x[:, :, 1::2, 1::2]) / 4.0


def residual_block(P, key, x, resample, dilation):
    """IN++ -> ELU -> conv1 -> IN++ -> ELU -> conv2(+pool) + shortcut. Reference: ResidualBlock,
    layers.py:401-456 (down & no dilation: conv2 and shortcut are ConvMeanPool 3x3 / 1x1;
    dilation: all three convs dilated, no pooling)."""
    d = 1 if dilation is None else dilation
    h = F.elu(instance_norm_plus(P, key + ".normalize1", x))
    h = _st(_conv(P, key + ".conv1", h, d))
    h = F.elu(instance_norm_plus(P, key + ".normalize2", h))
    pooled = resample == "down" and dilation is None
    if pooled:
        h = mean_pool2(_conv(P, key + ".conv2.conv", h, 1))
    else:
        h = _conv(P, key + ".conv2", h, d)
    if key + ".shortcut.conv.weight" in P:
        sc = _st(mean_pool2(_conv(P, key + ".shortcut.conv", x, 1, kernel=1)))
    elif key + ".shortcut.weight" in P:
        k = P[key + ".shortcut.weight"].shape[-1]
        sc = _st(_conv(P, key + ".shortcut", x, d if k == 3 else 1, kernel=k))
    else:
        sc = x
    if STREAM_ROUND is not None:
        return _st(sc + h), sc + h      # (stored stream, fp32 value): the decoder's skip operand is ELU of the latter
    return sc + h


def rcu(P, key, x, n_blocks, n_stages=2, exact=None):
    """Residual conv unit: per block r=x; (ELU, conv)*n_stages; x+=r. Reference: RCUBlock, layers.py:112-134.
    `exact` (STREAM_ROUND emulation only): the value of x before it was rounded into the 16-bit stream -- the kernels
    derive the ELU'd operand of the next convolution from the fp32 result, the residual from the stored stream.
    Returns x, or (x, exact) when `exact` is given."""
    xe = x if exact is None else exact
    for i in range(n_blocks):
        r = x
        h = xe
        for j in range(n_stages):
            h = _conv(P, f"{key}.{i + 1}_{j + 1}_conv", F.elu(h))
        xe = h + r
        x = _st(xe)
    return x if exact is None else (x, xe)


def crp(P, key, x, n_stages=2, exact=None):
    """Chained residual pooling with 5x5/s1 max-pool. Reference: CRPBlock, layers.py:62-83."""
    xe = x if exact is None else exact
    path = F.elu(xe)          # the pooled operand comes from the fp32 value, the residual ELU(x) from the stored stream
    x = F.elu(x)
    for i in range(n_stages):
        path = F.max_pool2d(path, kernel_size=5, stride=1, padding=2)
        path = _conv(P, f"{key}.convs.{i}", path)
        xe = path + x
        x = _st(xe)
    return x if exact is None else (x, xe)


def msf(P, key, xs, shape, want_exact=False):
    """sum_i bilinear_{align_corners}(conv_i(x_i)). Reference: MSFBlock, layers.py:165-184."""
    total = exact = None
    for i, xi in enumerate(xs):
        h = F.interpolate(_st(_conv(P, f"{key}.convs.{i}", xi)), size=shape, mode="bilinear", align_corners=True)
        exact = h if total is None else total + h
        total = _st(exact)
    return (total, exact) if want_exact else total


def refine(P, key, xs, shape, end=False):
    """adapt RCUs -> MSF (if >1 input) -> CRP -> output RCU (3 blocks when `end`).
    Reference: RefineBlock, layers.py:214-249.  xs: tensors, or (stream, exact) pairs under STREAM_ROUND emulation."""
    if STREAM_ROUND is None:
        hs = [rcu(P, f"{key}.adapt_convs.{i}", xi, 2) for i, xi in enumerate(xs)]
        h = msf(P, key + ".msf", hs, shape) if len(hs) > 1 else hs[0]
        h = crp(P, key + ".crp", h)
        return rcu(P, key + ".output_convs", h, 3 if end else 1)
    pairs = [xi if isinstance(xi, tuple) else (xi, xi) for xi in xs]
    hs = [rcu(P, f"{key}.adapt_convs.{i}", a, 2, exact=b) for i, (a, b) in enumerate(pairs)]
    if len(hs) > 1:
        h, he = msf(P, key + ".msf", [a for a, _ in hs], shape, want_exact=True)     # the MSF convolutions read the stored stream
    else:
        h, he = hs[0]
    h, he = crp(P, key + ".crp", h, exact=he)
    return rcu(P, key + ".output_convs", h, 3 if end else 1, exact=he)


# (stage name, [(resample, dilation) for the two blocks])
ENCODERS = {
    # Reference: NCSNv2.__init__, ncsnv2.py:29-63
    "NCSNv2": [("res1", [(None, None), (None, None)]), ("res2", [("down", None), (None, None)]),
               ("res3", [("down", 2), (None, 2)]), ("res4", [("down", 4), (None, 4)])],
    # Reference: NCSNv2Deeper.__init__, ncsnv2.py:122-154
    "NCSNv2Deeper": [("res1", [(None, None), (None, None)]), ("res2", [("down", None), (None, None)]),
                     ("res3", [("down", None), (None, None)]), ("res4", [("down", 2), (None, 2)]),
                     ("res5", [("down", 4), (None, 4)])],
    # Reference: NCSNv2Deepest.__init__, ncsnv2.py:215-255
    "NCSNv2Deepest": [("res1", [(None, None), (None, None)]), ("res2", [("down", None), (None, None)]),
                      ("res3", [("down", None), (None, None)]), ("res31", [("down", None), (None, None)]),
                      ("res4", [("down", 2), (None, 2)]), ("res5", [("down", 4), (None, 4)])],
}
# decoder: (refine name, skip stage index into the encoder list); deepest stage first
DECODERS = {
    "NCSNv2": ["refine1", "refine2", "refine3", "refine4"],                                   # ncsnv2.py:65-68,86-89
    "NCSNv2Deeper": ["refine1", "refine2", "refine3", "refine4", "refine5"],                   # ncsnv2.py:156-160,180-184
    "NCSNv2Deepest": ["refine1", "refine2", "refine31", "refine3", "refine4", "refine5"],      # ncsnv2.py:257-262,284-289
}


def score_forward(arch, P, x, labels, logit_transform=False, rescaled=False):
    """Reference: NCSNv2.forward (ncsnv2.py:70-101) / NCSNv2Deepest.forward (:269-299)."""
    h = 2 * x - 1.0 if (not logit_transform and not rescaled) else x
    h = _st(_conv(P, "begin_conv", h))
    feats = []
    for stage, blocks in ENCODERS[arch]:
        for i, (resample, dil) in enumerate(blocks):
            h = residual_block(P, f"{stage}.{i}", h, resample, dil)
            if isinstance(h, tuple):
                h, h_exact = h
        feats.append(h if STREAM_ROUND is None else (h, h_exact))
    names = DECODERS[arch]
    shape_of = lambda f: (f[0] if isinstance(f, tuple) else f).shape[2:]
    out = refine(P, names[0], [feats[-1]], shape_of(feats[-1]))
    for j, name in enumerate(names[1:], start=2):
        skip = feats[-j]
        out = refine(P, name, [skip, out], shape_of(skip), end=(j == len(names)))
    if isinstance(out, tuple):
        out = out[0]                    # the final InstanceNorm++ reads the stored stream
    out = F.elu(instance_norm_plus(P, "normalizer", out))
    out = _conv(P, "end_conv", out)
    sig = P["sigmas"][labels].view(x.shape[0], 1, 1, 1)
    return out / sig


def synth_state_dict(spec, seed, sigmas):
    """Deterministic random weights for a list of (key, shape): every tensor comes from its own
    torch.Generator seeded by (seed, position), so the reference model (via load_state_dict), this
    oracle and the CUDA path all see identical parameters without shipping them as fixtures.
    Conv weights ~ N(0, 0.36/fan_in) (about PyTorch's default scale, keeps the norm-free decoder in f16 range), biases ~ N(0, 0.05), alpha/gamma ~ N(1, 0.02) (the
    reference's own init, normalization.py:157-160), beta ~ N(0, 0.05)."""
    P = {}
    for idx, (key, shape) in enumerate(spec):
        if key == "sigmas":
            P[key] = sigmas.clone().float()
            continue
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        t = torch.randn(*shape, generator=g)
        if key.endswith(".weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if key == "conv_temporal_up.weight":       # ConvTranspose3d: (Cin, Cout, 1, 1, 4), two taps reach each output
                fan_in = shape[0] * 2
            t = t * (0.6 / fan_in ** 0.5)
        elif key.endswith(".alpha") or key.endswith(".gamma"):
            t = 1.0 + 0.02 * t
        else:
            t = 0.05 * t
        P[key] = t
    return P


# ------------------------------------------------------------------------------------------------ 3-D temporal prior
def _conv3d(P, key, x, dilation=1):
    """3x3x3 (pad = dilation) or 1x1x1 Conv3d. Reference: conv3x3 / dilated_conv3x3 / conv1x1, layers3d.py:29-60."""
    w = P[key + ".weight"]
    b = P.get(key + ".bias")
    k = w.shape[-1]
    if OPERAND_ROUND is not None and w.shape[0] > 1 and w.shape[1] > 1:
        x = x.to(OPERAND_ROUND).float()
        w = w.to(OPERAND_ROUND).float()
    return F.conv3d(x, w, b, stride=1, padding=dilation if k == 3 else 0, dilation=dilation if k == 3 else 1)


def instance_norm3d_plus(P, key, x, eps=1e-5):
    """Reference: InstanceNorm3dPlus.forward, normalization3d.py:164-185."""
    mu = x.mean(dim=(2, 3, 4))
    m_hat = (mu - mu.mean(dim=-1, keepdim=True)) / torch.sqrt(mu.var(dim=-1, keepdim=True) + eps)
    var = x.var(dim=(2, 3, 4), unbiased=False, keepdim=True)
    h = (x - mu[..., None, None, None]) / torch.sqrt(var + eps)
    h = h + (m_hat * P[key + ".alpha"])[..., None, None, None]
    out = P[key + ".gamma"].view(1, -1, 1, 1, 1) * h
    if key + ".beta" in P:
        out = out + P[key + ".beta"].view(1, -1, 1, 1, 1)
    return out


def residual_block3d(P, key, x, dilation):
    """Reference: ResidualBlock (3-D), layers3d.py:423-476; only the un-pooled variants NCSN3DShallow builds."""
    d = 1 if dilation is None else dilation
    h = F.elu(instance_norm3d_plus(P, key + ".normalize1", x))
    h = _conv3d(P, key + ".conv1", h, d)
    h = F.elu(instance_norm3d_plus(P, key + ".normalize2", h))
    h = _conv3d(P, key + ".conv2", h, d)
    sc = _conv3d(P, key + ".shortcut", x, d) if key + ".shortcut.weight" in P else x
    return sc + h


def rcu3d(P, key, x, n_blocks, n_stages=2):
    """Reference: RCUBlock (3-D), layers3d.py:113-135."""
    for i in range(n_blocks):
        r = x
        for j in range(n_stages):
            x = _conv3d(P, f"{key}.{i + 1}_{j + 1}_conv", F.elu(x))
        x = x + r
    return x


def crp3d(P, key, x, n_stages=2):
    """Reference: CRPBlock (3-D, MaxPool3d(5, 1, 2)), layers3d.py:63-84."""
    x = F.elu(x)
    path = x
    for i in range(n_stages):
        path = F.max_pool3d(path, kernel_size=5, stride=1, padding=2)
        path = _conv3d(P, f"{key}.convs.{i}", path)
        x = path + x
    return x


def refine3d(P, key, xs, shape, end=False):
    """Reference: RefineBlock / MSFBlock (3-D, trilinear align_corners), layers3d.py:166-187,219-255."""
    hs = [rcu3d(P, f"{key}.adapt_convs.{i}", xi, 2) for i, xi in enumerate(xs)]
    if len(hs) > 1:
        h = None
        for i, hi in enumerate(hs):
            t = F.interpolate(_conv3d(P, f"{key}.msf.convs.{i}", hi), size=shape, mode="trilinear", align_corners=True)
            h = t if h is None else h + t
    else:
        h = hs[0]
    h = crp3d(P, key + ".crp", h)
    return rcu3d(P, key + ".output_convs", h, 3 if end else 1)


def score_forward_3d_shallow(P, x, labels, logit_transform=False, rescaled=False):
    """x (B, 1, kx, ky, T). Reference: NCSN3DShallow.forward, ncsn3d.py:186-224."""
    h = 2 * x - 1.0 if (not logit_transform and not rescaled) else x
    out = _conv3d(P, "begin_conv", h)
    l1 = residual_block3d(P, "res1.1", residual_block3d(P, "res1.0", out, None), None)
    l2 = residual_block3d(P, "res3.1", residual_block3d(P, "res3.0", l1, 2), 2)
    wd, wu = P["conv_temporal_down.weight"], P["conv_temporal_up.weight"]
    xd = l2
    if OPERAND_ROUND is not None:
        xd, wd = xd.to(OPERAND_ROUND).float(), wd.to(OPERAND_ROUND).float()
    l3 = F.conv3d(xd, wd, P["conv_temporal_down.bias"], stride=(1, 1, 2), padding=(0, 0, 1))
    l4 = residual_block3d(P, "res4.1", residual_block3d(P, "res4.0", l3, 4), 4)
    r1 = refine3d(P, "refine1", [l4], l4.shape[2:])
    r2 = refine3d(P, "refine2", [l3, r1], l3.shape[2:])
    xu = r2
    if OPERAND_ROUND is not None:
        xu, wu = xu.to(OPERAND_ROUND).float(), wu.to(OPERAND_ROUND).float()
    r3 = F.conv_transpose3d(xu, wu, P["conv_temporal_up.bias"], stride=(1, 1, 2), padding=(0, 0, 1))
    o = refine3d(P, "refine3", [l1, r3], l1.shape[2:])
    o = F.elu(instance_norm3d_plus(P, "normalizer", o))
    o = _conv3d(P, "end_conv", o)
    return o / P["sigmas"][labels].view(x.shape[0], 1, 1, 1, 1)
