"""ORACLE (test infrastructure) — the fp32 oracle network evaluated with bf16 STORAGE emulated at exactly the points
where the B200 plan stores bf16 (activations, staged weights, activation gradients), everything else in fp32.

Why: on random weights / random labels the gap between ANY bf16 pipeline and the fp32 oracle is dominated by bf16
itself (stock torch autocast shows the same gap, see tools/parity_probe.py), so it cannot separate "bf16 noise" from
"wiring bug".  Against this emulation only accumulation order differs, so gradients must agree tightly; a wrong mask,
a missing residual or a dropped accumulation shows up as an O(1) error.

Rounding points (unet_b200/network.py):
  forward : network input; every conv weight; raw conv output before BatchNorm; every BN+ReLU / block-tail / decoder
            conv(+bias,+res,ReLU) output; blur(PixelShuffle) and relu(bn(skip)) when written into the concat buffer.
  backward: every stored activation gradient (the same tensors), and dlogits; where an activation has several
            consumers (skip features, block inputs with a convolutional shortcut) the plan accumulates their gradients
            one after the other in bf16 storage - `fork` reproduces that order and those roundings (the backward pass
            is so ill-conditioned on random-init networks that a single unmatched rounding moves deep gradients by tens
            of percent).
PARITY UNPINNED (see oracle/unet_oracle.py): this follows the same restated fastai graph.

Teacher forcing (`taps=`): random-init networks with batch statistics are chaotic - a 1e-7 relative perturbation in
front of the bf16 roundings (what a different fp32 accumulation order amounts to) moves deep-layer gradients by O(1)
(tests/test_cpu.py::test_teacher_forced_emulation measures it), so a free-running emulation cannot pin the wiring of
the deep layers.  With `taps = {rounding point name: tensor}` every named rounding point takes its FORWARD value from
the tap (the CUDA plan's stored activation) while gradients still flow through this graph: the backward pass is then a
linear map with exactly the plan's coefficients (same ReLU masks, same BatchNorm inputs), rounding-flip noise is no
longer amplified, and every parameter gradient must agree tightly.  `record=` collects the rounding points of a run.
Names are the plan's activation names (unet_b200/network.py `_act(..., name)`): `<conv>.raw`, `<conv>.out`,
`<block>.out`, `enc.bnrelu`, `layers.N.cat`, `input`, and `layers.5.conv2.2.{q,k,v,beta,o,out}` of the SelfAttention block.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .unet_oracle import DynamicUnetOracle, ResBlock, UnetBlock


class _Q(torch.autograd.Function):
    """round-to-bf16 in forward (storage of the activation) and in backward (storage of its gradient)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        if _Ctx.noise:                    # sensitivity probe: perturb in front of the gradient's rounding as well
            g = g * (1 + _Ctx.noise * torch.randn_like(g))
        return g.to(torch.bfloat16).to(torch.float32)


class _QGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _Fork(torch.autograd.Function):
    """A stored activation with several consumers: the plan's backward writes the first consumer's gradient (rounded
    to bf16), and every later consumer is added to the STORED value and rounded again (the dgrad epilogue reads the
    gradient tensor back as its residual).  Outputs are ordered as the plan's backward writes them."""

    @staticmethod
    def forward(ctx, x, n):
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *gs):
        acc = gs[0].to(torch.bfloat16).to(torch.float32)
        for g in gs[1:]:
            acc = (acc + g).to(torch.bfloat16).to(torch.float32)
        return acc, None


def fork(x, n):
    return _Fork.apply(x, n)


class _Ctx:
    """per-call state of emulated_forward: teacher-forcing taps, recording dict, noise level (sensitivity probe)"""
    taps = None
    record = None
    mismatch = None
    noise = 0.0


def q(x, name=None):
    if _Ctx.noise:
        x = x * (1 + _Ctx.noise * torch.randn_like(x))
    y = _Q.apply(x)
    if name is not None:
        if _Ctx.taps is not None and name in _Ctx.taps:
            t = _Ctx.taps[name].to(y.dtype)
            assert t.shape == y.shape, (name, tuple(t.shape), tuple(y.shape))
            if _Ctx.mismatch is not None:     # what THIS graph computes from the previous taps vs the tap itself
                _Ctx.mismatch[name] = ((t - y.detach()).abs().max() / t.abs().max().clamp_min(1e-30)).item()
            y = y + (t - y).detach()          # forward value: the tap; gradient: through this graph
        if _Ctx.record is not None:
            if y.requires_grad:
                y.retain_grad()           # after backward: the (unrounded) gradient arriving at this stored activation
            _Ctx.record[name] = y
    return y


def wq(w):
    """bf16-staged weight with a straight-through gradient to the fp32 master."""
    return w + (w.to(torch.bfloat16).to(torch.float32) - w).detach()


def _conv(x, conv, stride=None):
    return F.conv2d(x, wq(conv.weight), None, stride=conv.stride, padding=conv.padding)


def _bn(x, bn, training):
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)


def _conv_layer_bn(x, layer, training, act, name=None):
    """encoder ConvLayer: conv -> (stored raw, bf16) -> BN [-> ReLU]; returns the UNROUNDED BN output."""
    r = q(_conv(x, layer[0]), name and name + ".raw")
    z = _bn(r, layer[1], training)
    return F.relu(z) if act else z


def _has_idconv(blk: ResBlock) -> bool:
    return any(not isinstance(m, torch.nn.AvgPool2d) for m in blk.idpath)


def _res_block_bn(x, blk: ResBlock, training, name):
    """x: the block input, or (conv path input, idpath input) when the caller already forked it (skip features)."""
    n = len(blk.convpath)
    if isinstance(x, tuple):
        x, x_id = x
    elif _has_idconv(blk):
        x_id, x = fork(x, 2)      # the plan's backward writes the idpath gradient first, then adds the conv path's
    else:
        x_id = x                  # identity shortcut: dgrad + dOut in one epilogue, a single rounding
    h = x
    for j, layer in enumerate(blk.convpath):
        last = j == n - 1
        cn = f"{name}.convpath.{j}"
        z = _conv_layer_bn(h, layer, training, act=not last, name=cn)
        h = z if last else q(z, cn + ".out")
    idp = x_id
    for k, m in enumerate(blk.idpath):
        if isinstance(m, torch.nn.AvgPool2d):
            idp = m(idp)                       # fused into the 1x1 conv taps: not stored, not rounded
        else:
            idp = _conv_layer_bn(idp, m, training, act=False, name=f"{name}.idpath.{k}")
    return q(F.relu(h + idp), name + ".out")


def _conv_bias(x, layer, relu=True, res=None, name=None):
    y = _conv(x, layer[0]) + layer[0].bias.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    return q(F.relu(y), name and name + ".out") if relu else y


def _self_attention(x, sa, name, training):
    """fastai SelfAttention behind UnetBlock.conv2 with the plan's rounding points (unet_b200/network.py
    `_self_attention`): spectral-normed 1x1 query / key / value convolutions (one power iteration in training mode,
    sigma = u^T W v differentiated through W as torch.nn.utils.spectral_norm does), stored q / k / v, the logits
    S[i][j] = q_i . k_j (stored, but overwritten by the backward pass: not a tap), beta = softmax over i (stored; tap
    layout [N, j, i_h, i_w] = the plan's [image][i][j] tensor read as NHWC), o = sum_i beta_ij v_i, out = gamma o + x.
    The plan's backward writes x's gradient as (d out + W_q^T dQ) first, then adds W_k^T dK and W_v^T dV to the stored
    value: three consumers in that order."""
    N, C, H, W = x.shape
    n = H * W
    xa, xb, xc = fork(x, 3)

    def sn(layer):
        conv = layer[0]
        Wo = conv.weight_orig
        Wm = Wo.flatten(1)
        u, v = conv.weight_u.detach().clone(), conv.weight_v.detach().clone()
        if training:
            with torch.no_grad():
                v = F.normalize(Wm.t() @ u, dim=0, eps=1e-12)
                u = F.normalize(Wm @ v, dim=0, eps=1e-12)
        sigma = torch.dot(u, Wm @ v)
        return wq(Wo / sigma)

    proj = lambda t, layer, nm: q(F.conv1d(t.flatten(2), sn(layer)).view(N, -1, H, W), nm).flatten(2)
    f, g, h = proj(xa, sa.query, name + ".q"), proj(xb, sa.key, name + ".k"), proj(xc, sa.value, name + ".v")
    S = q(torch.bmm(f.transpose(1, 2), g))                                  # [N, i, j]
    beta = q(F.softmax(S, dim=1).permute(0, 2, 1).reshape(N, n, H, W), name + ".beta")
    beta = beta.reshape(N, n, n).permute(0, 2, 1)                           # back to [N, i, j]
    o = q(torch.bmm(h, beta).view(N, C, H, W), name + ".o")
    return q(sa.gamma * o + xa, name + ".out")


def emulated_forward(model: DynamicUnetOracle, x: torch.Tensor, training: bool = True, taps=None, record=None,
                     noise: float = 0.0, mismatch=None) -> torch.Tensor:
    """taps / record: see the module docstring (teacher forcing).  mismatch (dict, with taps): per tapped rounding point
    the max-norm relative difference between the tap and what this graph computes from the PREVIOUS taps - a per-layer
    forward check.  noise: relative Gaussian perturbation applied in front of every rounding (forward and backward) - a
    probe of the conditioning, standing in for the difference between two fp32 accumulation orders."""
    _Ctx.taps, _Ctx.record, _Ctx.noise, _Ctx.mismatch = taps, record, noise, mismatch
    try:
        return _emulated_forward(model, x, training)
    finally:
        _Ctx.taps, _Ctx.record, _Ctx.noise, _Ctx.mismatch = None, None, 0.0, None


def _emulated_forward(model: DynamicUnetOracle, x: torch.Tensor, training: bool) -> torch.Tensor:
    L = model.layers
    x0 = q(x, "input")
    feats = {}
    h = x0
    enc = L[0]
    for i, child in enumerate(enc):
        if i < 3:
            h = q(_conv_layer_bn(h, child, training, act=True, name=f"layers.0.{i}"), f"layers.0.{i}.out")
        elif i == 3:
            h = child(h)
        else:
            for b, blk in enumerate(child):
                h = _res_block_bn(h, blk, training, f"layers.0.{i}.{b}")
        if i in model.SKIP_IDXS:
            # a skip feature: the decoder's gradient is written first, the encoder's own consumers are added to it
            nxt = enc[i + 1]
            if isinstance(nxt, torch.nn.Sequential) and isinstance(nxt[0], ResBlock) and _has_idconv(nxt[0]):
                feats[i], h_id, h_conv = fork(h, 3)
                h = (h_conv, h_id)
            else:
                feats[i], h = fork(h, 2)
    h = q(F.relu(_bn(h, L[1], training)), "enc.bnrelu")
    for k, layer in enumerate(L[3]):
        h = _conv_bias(h, layer, name=f"layers.3.{k}")
    for j, idx in enumerate(model.SKIP_IDXS):
        ub: UnetBlock = L[4 + j]
        s = feats[idx]
        p = _conv_bias(h, ub.shuf[0], name=f"layers.{4 + j}.shuf.0")
        up = ub.shuf[3](ub.shuf[2](ub.shuf[1](p)))
        if s.shape[-2:] != up.shape[-2:]:
            up = F.interpolate(up, s.shape[-2:], mode="nearest")
        cat = q(F.relu(torch.cat([up, _bn(s, ub.bn, training)], dim=1)), f"layers.{4 + j}.cat")
        h = _conv_bias(_conv_bias(cat, ub.conv1, name=f"layers.{4 + j}.conv1"), ub.conv2, name=f"layers.{4 + j}.conv2")
        if len(ub.conv2) > 2:                  # ConvLayer(..., xtra=SelfAttention): module "2" behind conv + ReLU
            h = _self_attention(h, ub.conv2[2], f"layers.{4 + j}.conv2.2", training)
    p8 = _conv_bias(h, L[8][0], name="layers.8.0")
    up8 = L[8][1](p8)
    if up8.shape[-2:] != x0.shape[-2:]:
        up8 = F.interpolate(up8, x0.shape[-2:], mode="nearest")
    cat = torch.cat([up8, x0], dim=1)
    rb = L[11]
    a1 = _conv_bias(cat, rb.convpath[0], name="layers.11.convpath.0")
    a2 = _conv_bias(a1, rb.convpath[1], relu=True, res=cat, name="layers.11.convpath.1")
    logits = _conv(a2, L[12][0]) + L[12][0].bias.view(1, -1, 1, 1)
    return _QGrad.apply(logits)
