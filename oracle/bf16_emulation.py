"""ORACLE (test infrastructure) — the fp32 oracle network evaluated with bf16 STORAGE emulated at exactly the points
where the B200 plan stores bf16 (activations, staged weights, activation gradients), everything else in fp32.

Why: on random weights / random labels the gap between ANY bf16 pipeline and the fp32 oracle is dominated by bf16
itself (stock torch autocast shows the same gap, see tools/parity_probe.py), so it cannot separate "bf16 noise" from
"wiring bug".  Against this emulation only accumulation order differs, so gradients must agree tightly; a wrong mask,
a missing residual or a dropped accumulation shows up as an O(1) error.

Rounding points (unet_b200/network.py):
  forward : network input; every conv weight; raw conv output before BatchNorm; every BN+ReLU / block-tail / decoder
            conv(+bias,+res,ReLU) output; blur(PixelShuffle) and relu(bn(skip)) when written into the concat buffer.
  backward: every stored activation gradient (the same tensors), and dlogits.
PARITY UNPINNED (see oracle/unet_oracle.py): this follows the same restated fastai graph.
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
        return g.to(torch.bfloat16).to(torch.float32)


class _QGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def q(x):
    return _Q.apply(x)


def wq(w):
    """bf16-staged weight with a straight-through gradient to the fp32 master."""
    return w + (w.to(torch.bfloat16).to(torch.float32) - w).detach()


def _conv(x, conv, stride=None):
    return F.conv2d(x, wq(conv.weight), None, stride=conv.stride, padding=conv.padding)


def _bn(x, bn, training):
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)


def _conv_layer_bn(x, layer, training, act):
    """encoder ConvLayer: conv -> (stored raw, bf16) -> BN [-> ReLU]; returns the UNROUNDED BN output."""
    r = q(_conv(x, layer[0]))
    z = _bn(r, layer[1], training)
    return F.relu(z) if act else z


def _res_block_bn(x, blk: ResBlock, training):
    n = len(blk.convpath)
    h = x
    for j, layer in enumerate(blk.convpath):
        last = j == n - 1
        z = _conv_layer_bn(h, layer, training, act=not last)
        h = z if last else q(z)
    idp = x
    for m in blk.idpath:
        if isinstance(m, torch.nn.AvgPool2d):
            idp = m(idp)                       # fused into the 1x1 conv taps: not stored, not rounded
        else:
            idp = _conv_layer_bn(idp, m, training, act=False)
    return q(F.relu(h + idp))


def _conv_bias(x, layer, relu=True, res=None):
    y = _conv(x, layer[0]) + layer[0].bias.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    return q(F.relu(y)) if relu else y


def emulated_forward(model: DynamicUnetOracle, x: torch.Tensor, training: bool = True) -> torch.Tensor:
    L = model.layers
    x0 = q(x)
    feats = {}
    h = x0
    enc = L[0]
    for i, child in enumerate(enc):
        if i < 3:
            h = q(_conv_layer_bn(h, child, training, act=True))
        elif i == 3:
            h = child(h)
        else:
            for blk in child:
                h = _res_block_bn(h, blk, training)
        if i in model.SKIP_IDXS:
            feats[i] = h
    h = q(F.relu(_bn(h, L[1], training)))
    for layer in L[3]:
        h = _conv_bias(h, layer)
    for j, idx in enumerate(model.SKIP_IDXS):
        ub: UnetBlock = L[4 + j]
        s = feats[idx]
        p = _conv_bias(h, ub.shuf[0])
        up = ub.shuf[3](ub.shuf[2](ub.shuf[1](p)))
        if s.shape[-2:] != up.shape[-2:]:
            up = F.interpolate(up, s.shape[-2:], mode="nearest")
        cat = q(F.relu(torch.cat([up, _bn(s, ub.bn, training)], dim=1)))
        h = _conv_bias(_conv_bias(cat, ub.conv1), ub.conv2)
    p8 = _conv_bias(h, L[8][0])
    up8 = L[8][1](p8)
    if up8.shape[-2:] != x0.shape[-2:]:
        up8 = F.interpolate(up8, x0.shape[-2:], mode="nearest")
    cat = torch.cat([up8, x0], dim=1)
    rb = L[11]
    a1 = _conv_bias(cat, rb.convpath[0])
    a2 = _conv_bias(a1, rb.convpath[1], relu=True, res=cat)
    logits = _conv(a2, L[12][0]) + L[12][0].bias.view(1, -1, 1, 1)
    return _QGrad.apply(logits)
