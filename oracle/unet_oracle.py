"""ORACLE (test infrastructure, never the product path) — plain-PyTorch fp32 restatement of the network the reference
asks fastai to build.

PARITY UNPINNED: the reference (LUP-LuftbildUmweltPlanung/UNet) has no tests or golden vectors, and the arithmetic lives
in un-vendored third-party packages that are absent from /root/reference and from this image:
    fastai==2.5.1 (environment/requirements.txt:4), torch==1.9.1 (requirements.txt:11).
This file restates fastai 2.5.1's published module graph (fastai/vision/models/xresnet.py, fastai/vision/models/unet.py,
fastai/layers.py) for exactly the flags the reference passes, anchored on the reference's own call sites:

    create_body(arch, pretrained, cut=None)                               train.py:128
    body[0][0] = nn.Conv2d(n_in, 32, 3, stride 2, pad 1, bias=None)       train.py:130-135
    DynamicUnet(body, n_out, img_size, blur=True, blur_final=True, self_attention=..., y_range=None,
                norm_type=NormType (the Enum class => no decoder norm, conv bias on), last_cross=True, bottle=False)
                                                                          train.py:141-144
    CrossEntropyLossFlat(axis=1) with .func.weight = class weights         train.py:187,195,211
    fastai Adam / fit_one_cycle                                           train.py:218,246-250

Module and parameter names follow fastai's so that `state_dict()` keys are the ones a reference-trained checkpoint has
(`layers.0.4.0.convpath.0.0.weight`, `layers.4.shuf.0.0.bias`, ...; SURVEY.md 8(a) key map).

Pieces that DO have an independent implementation in this image are pinned against it (tests/test_cpu.py): the residual
stages against torchvision's ResNet stages (numerically, same weights), fastai_adam_step against torch.optim.AdamW,
dice_multi against scikit-learn's macro F1.  The fastai-specific wiring stays unpinned.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

ARCHS: Dict[str, Tuple[int, Sequence[int]]] = {
    # name -> (expansion, blocks per stage)       fastai xresnet.py: xresnet18/34/50/101
    "xresnet18": (1, (2, 2, 2, 2)),
    "xresnet34": (1, (3, 4, 6, 3)),
    "xresnet50": (4, (3, 4, 6, 3)),
    "xresnet101": (4, (3, 4, 23, 3)),
}


def conv_layer(ni: int, nf: int, ks: int = 3, stride: int = 1, bn: bool = True, act: bool = True,
               zero_bn: bool = False) -> nn.Sequential:
    """fastai ConvLayer: Conv2d(bias = not bn, padding=(ks-1)//2) -> [BatchNorm2d] -> [ReLU]   (layers.py ConvLayer).
    BatchNorm init: weight 1 (0 for BatchZero), bias 1e-3 (layers.py _get_norm)."""
    layers: List[nn.Module] = [nn.Conv2d(ni, nf, ks, stride=stride, padding=(ks - 1) // 2, bias=not bn)]
    if bn:
        b = nn.BatchNorm2d(nf)
        b.bias.data.fill_(1e-3)
        b.weight.data.fill_(0.0 if zero_bn else 1.0)
        layers.append(b)
    if act:
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class ResBlock(nn.Module):
    """fastai layers.ResBlock for expansion 1 / 4, with (encoder) or without (decoder, norm_type junk) BatchNorm."""

    def __init__(self, expansion: int, ni: int, nf: int, stride: int = 1, bn: bool = True):
        super().__init__()
        nh = nf
        nf, ni = nf * expansion, ni * expansion
        if expansion == 1:
            convpath = [conv_layer(ni, nh, 3, stride, bn=bn), conv_layer(nh, nf, 3, bn=bn, act=False, zero_bn=True)]
        else:
            convpath = [conv_layer(ni, nh, 1, bn=bn), conv_layer(nh, nh, 3, stride, bn=bn),
                        conv_layer(nh, nf, 1, bn=bn, act=False, zero_bn=True)]
        self.convpath = nn.Sequential(*convpath)
        idpath: List[nn.Module] = []
        if ni != nf:
            idpath.append(conv_layer(ni, nf, 1, bn=True, act=False))  # idpath ConvLayer keeps the default Batch norm
        if stride != 1:
            idpath.insert(0, nn.AvgPool2d(2, ceil_mode=True))  # pool_first=True
        self.idpath = nn.Sequential(*idpath)
        self.act = nn.ReLU()

    def forward(self, x):
        return self.act(self.convpath(x) + self.idpath(x))


def xresnet_body(arch: str, n_in: int) -> nn.Sequential:
    """children 0..7 of fastai XResNet (create_body cut=None cuts before AdaptiveAvgPool), first conv taking n_in."""
    expansion, layers = ARCHS[arch]
    sizes = [n_in, 32, 32, 64]
    stem = [conv_layer(sizes[i], sizes[i + 1], 3, stride=2 if i == 0 else 1) for i in range(3)]
    block_szs = [64 // expansion, 64, 128, 256, 512]
    stages = []
    for i, l in enumerate(layers):
        ni, nf = block_szs[i], block_szs[i + 1]
        stride = 1 if i == 0 else 2
        stages.append(nn.Sequential(*[ResBlock(expansion, ni if b == 0 else nf, nf, stride if b == 0 else 1)
                                      for b in range(l)]))
    return nn.Sequential(*stem, nn.MaxPool2d(3, 2, padding=1), *stages)


def pixel_shuffle_icnr(ni: int, nf: int, blur: bool) -> nn.Sequential:
    layers: List[nn.Module] = [conv_layer(ni, nf * 4, 1, bn=False, act=True), nn.PixelShuffle(2)]
    if blur:
        layers += [nn.ReplicationPad2d((1, 0, 1, 0)), nn.AvgPool2d(2, stride=1)]
    return nn.Sequential(*layers)


class SelfAttention(nn.Module):
    """fastai layers.SelfAttention(n_channels) [fastai-mem, 2.5.1]: query / key / value are
    ConvLayer(ndim=1, ks=1, norm_type=NormType.Spectral, act_cls=None, bias=False) = Sequential(spectral_norm(Conv1d));
    beta = softmax(bmm(f^T, g), dim=1); o = gamma * bmm(h, beta) + x; gamma is initialised to 0.
    Own parameter (gamma) precedes the children in named_parameters(), as in fastai."""

    def __init__(self, n_channels: int):
        super().__init__()
        self.gamma = nn.Parameter(torch.tensor([0.0]))
        mk = lambda co: nn.Sequential(nn.utils.spectral_norm(nn.Conv1d(n_channels, co, 1, bias=False)))
        self.query, self.key, self.value = mk(n_channels // 8), mk(n_channels // 8), mk(n_channels)

    def forward(self, x):
        size = x.size()
        x = x.view(*size[:2], -1)
        f, g, h = self.query(x), self.key(x), self.value(x)
        beta = F.softmax(torch.bmm(f.transpose(1, 2), g), dim=1)
        o = self.gamma * torch.bmm(h, beta) + x
        return o.view(*size).contiguous()


class UnetBlock(nn.Module):
    def __init__(self, up_in_c: int, x_in_c: int, final_div: bool, blur: bool, self_attention: bool = False):
        super().__init__()
        self.shuf = pixel_shuffle_icnr(up_in_c, up_in_c // 2, blur)
        self.bn = nn.BatchNorm2d(x_in_c)
        self.bn.bias.data.fill_(1e-3)
        ni = up_in_c // 2 + x_in_c
        nf = ni if final_div else ni // 2
        self.conv1 = conv_layer(ni, nf, 3, bn=False)
        self.conv2 = conv_layer(nf, nf, 3, bn=False)
        if self_attention:      # ConvLayer(nf, nf, xtra=SelfAttention(nf)): appended after the activation -> conv2.2
            self.conv2.add_module("2", SelfAttention(nf))
        self.relu = nn.ReLU()
        self.nf = nf

    def forward(self, up_in, s):
        up_out = self.shuf(up_in)
        if s.shape[-2:] != up_out.shape[-2:]:
            up_out = F.interpolate(up_out, s.shape[-2:], mode="nearest")
        cat_x = self.relu(torch.cat([up_out, self.bn(s)], dim=1))
        return self.conv2(self.conv1(cat_x))


class _Identity(nn.Module):
    def forward(self, x):
        return x


class DynamicUnetOracle(nn.Module):
    """DynamicUnet(body, n_out, blur=True, blur_final=True, last_cross=True, bottle=False, no decoder norm)."""

    SKIP_IDXS = (6, 5, 4, 2)  # encoder children whose successor halves the resolution, deepest first

    def __init__(self, arch: str = "xresnet34", n_in: int = 4, n_out: int = 2, self_attention: bool = False):
        super().__init__()
        expansion, _ = ARCHS[arch]
        enc = xresnet_body(arch, n_in)
        widths = {2: 64, 4: 64 * expansion, 5: 128 * expansion, 6: 256 * expansion, 7: 512 * expansion}
        ni = widths[7]
        bn = nn.BatchNorm2d(ni)
        bn.bias.data.fill_(1e-3)
        middle = nn.Sequential(conv_layer(ni, ni * 2, 3, bn=False), conv_layer(ni * 2, ni, 3, bn=False))
        layers: List[nn.Module] = [enc, bn, nn.ReLU(), middle]
        c = ni
        for i, idx in enumerate(self.SKIP_IDXS):
            not_final = i != len(self.SKIP_IDXS) - 1
            # fastai unet.py: sa = self_attention and (i == len(sz_chg_idxs) - 3)  -> the second UnetBlock (layers.5)
            sa = self_attention and i == len(self.SKIP_IDXS) - 3
            blk = UnetBlock(c, widths[idx], final_div=not_final, blur=True, self_attention=sa)
            layers.append(blk)
            c = blk.nf
        layers.append(pixel_shuffle_icnr(c, c, blur=False))          # layers.8
        layers.append(_Identity())                                    # layers.9  ResizeToOrig (no-op at 128/256/512)
        layers.append(_Identity())                                    # layers.10 MergeLayer(dense=True): cat([x, input])
        c += n_in
        layers.append(ResBlock(1, c, c, bn=False))                    # layers.11
        layers.append(conv_layer(c, n_out, 1, bn=False, act=False))   # layers.12
        layers.append(_Identity())                                    # layers.13 ToTensorBase
        self.layers = nn.ModuleList(layers)
        self.arch, self.n_in, self.n_out = arch, n_in, n_out

    def forward(self, x):
        inp = x
        enc = self.layers[0]
        feats = {}
        for i, child in enumerate(enc):
            x = child(x)
            if i in self.SKIP_IDXS:
                feats[i] = x
        x = self.layers[3](self.layers[2](self.layers[1](x)))
        for j, idx in enumerate(self.SKIP_IDXS):
            x = self.layers[4 + j](x, feats[idx])
        x = self.layers[8](x)
        if x.shape[-2:] != inp.shape[-2:]:
            x = F.interpolate(x, inp.shape[-2:], mode="nearest")
        x = torch.cat([x, inp], dim=1)
        x = self.layers[11](x)
        return self.layers[12](x)


def init_like_fastai(model: DynamicUnetOracle, seed: int = 0) -> None:
    """kaiming_normal on every conv (init_cnn / apply_init), bias 0 — irrelevant for parity (weights are injected)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
            m.weight.data.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
            if m.bias is not None:
                m.bias.data.zero_()


def randomize_bn(model: nn.Module, seed: int = 1) -> None:
    """Stock init has gamma=0 on the last BN of every ResBlock (BatchZero), which would make parity tests vacuous
    (SURVEY.md 7 'Zero-init BN'): draw gamma ~ U(0.5,1.5), beta ~ N(0,0.1), running stats away from (0,1)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    # decoder biases away from zero as well
    for m in model.modules():
        if isinstance(m, nn.Conv2d) and m.bias is not None:
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)


def make_oracle(arch: str = "xresnet34", n_in: int = 4, n_out: int = 2, seed: int = 0,
                self_attention: bool = False) -> DynamicUnetOracle:
    torch.manual_seed(seed)          # spectral_norm draws its u / v vectors from the global generator at construction
    m = DynamicUnetOracle(arch, n_in, n_out, self_attention)
    init_like_fastai(m, seed)
    randomize_bn(m, seed + 1)
    if self_attention:
        g = torch.Generator().manual_seed(seed + 2)
        for mod in m.modules():
            if isinstance(mod, SelfAttention):
                mod.gamma.data.fill_(0.5)     # fastai starts at 0 (identity): move it so that parity tests see the block
                for c in (mod.query[0], mod.key[0], mod.value[0]):
                    fan_in = c.in_channels
                    c.weight_orig.data.copy_(torch.randn(c.weight_orig.shape, generator=g) * math.sqrt(2.0 / fan_in))
    return m


# ------------------------------------------------------------------------------------------------ loss / metric / opt
def weighted_ce(logits: torch.Tensor, target: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """CrossEntropyLossFlat(axis=1): logits [B,C,H,W] -> [-1,C], target [-1]; weighted mean (fastai losses.py)."""
    c = logits.shape[1]
    return F.cross_entropy(logits.permute(0, 2, 3, 1).reshape(-1, c), target.reshape(-1).long(), weight=weight)


def dice_multi(pred_classes: torch.Tensor, target: torch.Tensor, n_classes: int) -> float:
    """fastai DiceMulti (metrics.py): per-class 2*I/(P+T) accumulated over the set, nan-mean over classes
    (a class absent from both prediction and target is ignored)."""
    vals = []
    for c in range(n_classes):
        p, t = pred_classes == c, target == c
        inter, union = (p & t).sum().item(), p.sum().item() + t.sum().item()
        vals.append(2.0 * inter / union if union > 0 else float("nan"))
    vals = [v for v in vals if not math.isnan(v)]
    return float(sum(vals) / len(vals)) if vals else float("nan")


def sgd_step(params, lr: float) -> None:
    """BASELINE config 1: plain SGD."""
    with torch.no_grad():
        for p in params:
            if p.grad is not None:
                p.add_(p.grad, alpha=-lr)


def fastai_adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
                     mom: float = 0.9, sqr_mom: float = 0.99, eps: float = 1e-5, wd: float = 0.01) -> None:
    """fastai optimizer.Adam with decouple_wd=True: weight_decay -> average_grad(dampening) -> average_sqr_grad ->
    step_stat -> adam_step (eps added after the sqrt, debiased by 1-mom**step / 1-sqr_mom**step)."""
    if wd != 0:
        p.mul_(1 - lr * wd)
    m.mul_(mom).add_(g, alpha=1 - mom)
    v.mul_(sqr_mom).addcmul_(g, g, value=1 - sqr_mom)
    debias1 = 1 - mom ** step
    debias2 = 1 - sqr_mom ** step
    p.addcdiv_(m, (v / debias2).sqrt() + eps, value=-lr / debias1)


def one_cycle_lr(pct: float, lr_max: float, div: float = 25.0, div_final: float = 1e5, pct_start: float = 0.25,
                 moms=(0.95, 0.85, 0.95)) -> Tuple[float, float]:
    """fastai fit_one_cycle: combined_cos(pct_start, lr_max/div, lr_max, lr_max/div_final), same for momentum."""
    def cos(a, b, p):
        return a + (1 + math.cos(math.pi * (1 - p))) * (b - a) / 2

    if pct < pct_start:
        q = pct / pct_start
        return cos(lr_max / div, lr_max, q), cos(moms[0], moms[1], q)
    q = (pct - pct_start) / (1 - pct_start)
    return cos(lr_max, lr_max / div_final, q), cos(moms[1], moms[2], q)


def param_groups(model: DynamicUnetOracle) -> List[List[str]]:
    """fastai _xresnet_split (mirrored at reference train.py:78-80): [body[:3], body[3:], decoder]."""
    g: List[List[str]] = [[], [], []]
    for name, _ in model.named_parameters():
        parts = name.split(".")
        if parts[1] == "0":
            g[0 if int(parts[2]) < 3 else 1].append(name)
        else:
            g[2].append(name)
    return g


def count_conv_flops(model: DynamicUnetOracle, size: int) -> int:
    """forward conv FLOPs per tile (2 FLOP per MAC, unpadded) — SURVEY.md 8(a)/(d) numerators."""
    total = 0
    hooks = []

    def hook(m, inp, out):
        nonlocal total
        total += 2 * out.numel() // out.shape[0] * m.in_channels * m.kernel_size[0] * m.kernel_size[1]

    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            hooks.append(m.register_forward_hook(hook))
    model.eval()
    with torch.no_grad():
        model(torch.zeros(1, model.n_in, size, size))
    for h in hooks:
        h.remove()
    return total
