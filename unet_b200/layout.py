"""Architecture description and parameter layout of the xresnet-DynamicUnet (CPU-only, no CUDA needed).

Mirrors what `unet_learner_MS` asks fastai to build (reference train.py:98-160; module graph in SURVEY.md 8(a)):
xresnet{18,34,50,101} body with an `n_in`-band first conv, `DynamicUnet(blur=True, blur_final=True, last_cross=True,
bottle=False)` without decoder norm.  Parameter names are fastai's state_dict keys, in `named_parameters()` order, so
that reference-trained weights load and the three fastai parameter groups (train.py:78-80) are contiguous ranges of the
flat fp32 parameter buffer.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

ARCHS: Dict[str, Tuple[int, Sequence[int]]] = {
    "xresnet18": (1, (2, 2, 2, 2)),
    "xresnet34": (1, (3, 4, 6, 3)),
    "xresnet50": (4, (3, 4, 6, 3)),
    "xresnet101": (4, (3, 4, 23, 3)),
}


@dataclass
class ConvSpec:
    name: str            # fastai prefix of the ConvLayer, e.g. "layers.0.4.0.convpath.0"
    ni: int
    nf: int
    ks: int
    stride: int = 1
    bn: bool = False     # conv -> BatchNorm (no conv bias) as in the encoder
    act: bool = True
    pool: bool = False   # AvgPool2d(2, ceil_mode=True) in front (ResBlock idpath with stride 2)
    shuffle: bool = False  # 1x1 conv of a PixelShuffle_ICNR: GEMM rows are permuted to (i, j, c) order
    sn: bool = False     # spectral-normed Conv1d(ks=1, bias=False) of SelfAttention: weight_orig [nf, ni, 1] + u / v buffers

    @property
    def wname(self) -> str:
        return self.name + (".0.weight_orig" if self.sn else ".0.weight")

    @property
    def bname(self) -> Optional[str]:
        return None if (self.bn or self.sn) else self.name + ".0.bias"

    @property
    def bn_prefix(self) -> Optional[str]:
        return self.name + ".1" if self.bn else None


@dataclass
class BlockSpec:
    name: str
    convpath: List[ConvSpec]
    idconv: Optional[ConvSpec]
    stride: int
    ni: int
    nf: int


@dataclass
class SASpec:
    """fastai SelfAttention(n_channels) appended to UnetBlock.conv2 (`conv2.2`): gamma + query / key / value."""
    name: str            # "layers.5.conv2.2"
    c: int
    query: ConvSpec
    key: ConvSpec
    value: ConvSpec

    @property
    def gamma(self) -> str:
        return self.name + ".gamma"

    def convs(self) -> List[ConvSpec]:
        return [self.query, self.key, self.value]


@dataclass
class UnetBlockSpec:
    name: str            # "layers.4" ...
    shuf: ConvSpec
    bn_prefix: str
    conv1: ConvSpec
    conv2: ConvSpec
    up_in_c: int
    x_in_c: int
    cu: int              # channels after the shuffle (up_in_c // 2)
    skip_child: int      # encoder child whose output is concatenated
    sa: Optional[SASpec] = None


@dataclass
class NetSpec:
    arch: str
    n_in: int
    n_out: int
    stem: List[ConvSpec]
    stages: List[List[BlockSpec]]
    enc_out_c: int
    post_bn: str
    middle: List[ConvSpec]
    unet: List[UnetBlockSpec]
    final_shuf: ConvSpec
    final_res: List[ConvSpec]
    head: ConvSpec
    skip_widths: Dict[int, int] = field(default_factory=dict)

    def convs(self) -> List[ConvSpec]:
        out = list(self.stem)
        for st in self.stages:
            for b in st:
                out += b.convpath
                if b.idconv is not None:
                    out.append(b.idconv)
        out += self.middle
        for u in self.unet:
            out += [u.shuf, u.conv1, u.conv2]
            if u.sa is not None:
                out += u.sa.convs()
        out += [self.final_shuf] + self.final_res + [self.head]
        return out


def build_spec(arch: str = "xresnet34", n_in: int = 4, n_out: int = 2, self_attention: bool = False) -> NetSpec:
    if arch not in ARCHS:
        raise ValueError(f"unsupported architecture {arch!r}; choose from {sorted(ARCHS)}")
    expansion, layers = ARCHS[arch]
    sizes = [n_in, 32, 32, 64]
    stem = [ConvSpec(f"layers.0.{i}", sizes[i], sizes[i + 1], 3, 2 if i == 0 else 1, bn=True) for i in range(3)]
    block_szs = [64 // expansion, 64, 128, 256, 512]
    stages: List[List[BlockSpec]] = []
    for si, nblocks in enumerate(layers):
        blocks = []
        for b in range(nblocks):
            ni0 = block_szs[si] if b == 0 else block_szs[si + 1]
            nf0 = block_szs[si + 1]
            stride = (1 if si == 0 else 2) if b == 0 else 1
            ni, nf, nh = ni0 * expansion, nf0 * expansion, nf0
            pre = f"layers.0.{4 + si}.{b}"
            if expansion == 1:
                cp = [ConvSpec(f"{pre}.convpath.0", ni, nh, 3, stride, bn=True),
                      ConvSpec(f"{pre}.convpath.1", nh, nf, 3, 1, bn=True, act=False)]
            else:
                cp = [ConvSpec(f"{pre}.convpath.0", ni, nh, 1, 1, bn=True),
                      ConvSpec(f"{pre}.convpath.1", nh, nh, 3, stride, bn=True),
                      ConvSpec(f"{pre}.convpath.2", nh, nf, 1, 1, bn=True, act=False)]
            idc = None
            if ni != nf:
                k = 1 if stride != 1 else 0   # idpath = [AvgPool, ConvLayer] when strided, else [ConvLayer]
                idc = ConvSpec(f"{pre}.idpath.{k}", ni, nf, 1, 1, bn=True, act=False, pool=stride != 1)
            blocks.append(BlockSpec(pre, cp, idc, stride, ni, nf))
        stages.append(blocks)
    widths = {2: 64, 4: 64 * expansion, 5: 128 * expansion, 6: 256 * expansion, 7: 512 * expansion}
    c = widths[7]
    middle = [ConvSpec("layers.3.0", c, 2 * c, 3), ConvSpec("layers.3.1", 2 * c, c, 3)]
    unet = []
    for j, idx in enumerate((6, 5, 4, 2)):
        not_final = j != 3
        up_in_c, x_in_c = c, widths[idx]
        cu = up_in_c // 2
        ni = cu + x_in_c
        nf = ni if not_final else ni // 2
        pre = f"layers.{4 + j}"
        unet.append(UnetBlockSpec(pre, ConvSpec(f"{pre}.shuf.0", up_in_c, 4 * cu, 1, shuffle=True), f"{pre}.bn",
                                  ConvSpec(f"{pre}.conv1", ni, nf, 3), ConvSpec(f"{pre}.conv2", nf, nf, 3),
                                  up_in_c, x_in_c, cu, idx))
        if self_attention and j == 1:      # fastai unet.py: sa = self_attention and (i == len(sz_chg_idxs) - 3)
            sp = f"{pre}.conv2.2"
            unet[-1].sa = SASpec(sp, nf, ConvSpec(f"{sp}.query", nf, nf // 8, 1, act=False, sn=True),
                                 ConvSpec(f"{sp}.key", nf, nf // 8, 1, act=False, sn=True),
                                 ConvSpec(f"{sp}.value", nf, nf, 1, act=False, sn=True))
        c = nf
    final_shuf = ConvSpec("layers.8.0", c, 4 * c, 1, shuffle=True)
    cc = c + n_in
    final_res = [ConvSpec("layers.11.convpath.0", cc, cc, 3), ConvSpec("layers.11.convpath.1", cc, cc, 3, act=False)]
    head = ConvSpec("layers.12", cc, n_out, 1, act=False)
    return NetSpec(arch, n_in, n_out, stem, stages, widths[7], "layers.1", middle, unet, final_shuf, final_res, head,
                   widths)


@dataclass
class ParamEntry:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int
    group: int       # fastai splitter group: 0 = body[:3], 1 = body[3:], 2 = decoder
    decay: bool      # weight decay applies (conv weights only; wd_bn_bias=False, train.py:102)


class ParamLayout:
    """Flat fp32 parameter buffer layout: name -> (offset, shape), in fastai named_parameters() order."""

    def __init__(self, spec: NetSpec):
        self.spec = spec
        self.entries: List[ParamEntry] = []
        self.by_name: Dict[str, ParamEntry] = {}
        self.buffers: List[Tuple[str, int]] = []   # BN running stats: (prefix, C)
        self.sn_buffers: List[Tuple[str, int, int]] = []   # spectral norm u / v vectors: (conv prefix, Cout, Cin)
        self._off = 0

        def group_of(name: str) -> int:
            parts = name.split(".")
            if parts[1] == "0":
                return 0 if int(parts[2]) < 3 else 1
            return 2

        def add(name: str, shape: Tuple[int, ...], decay: bool):
            n = 1
            for s in shape:
                n *= s
            # keep every tensor 16-byte aligned inside the flat buffer
            off = (self._off + 3) // 4 * 4
            e = ParamEntry(name, shape, off, n, group_of(name), decay)
            self.entries.append(e)
            self.by_name[name] = e
            self._off = off + n

        def add_conv(cs: ConvSpec):
            if cs.sn:
                add(cs.wname, (cs.nf, cs.ni, 1), True)
                self.sn_buffers.append((cs.name + ".0", cs.nf, cs.ni))
                return
            add(cs.wname, (cs.nf, cs.ni, cs.ks, cs.ks), True)
            if cs.bn:
                add_bn(cs.bn_prefix, cs.nf)
            else:
                add(cs.bname, (cs.nf,), False)

        def add_bn(prefix: str, c: int):
            add(prefix + ".weight", (c,), False)
            add(prefix + ".bias", (c,), False)
            self.buffers.append((prefix, c))

        for cs in spec.stem:
            add_conv(cs)
        for st in spec.stages:
            for b in st:
                for cs in b.convpath:
                    add_conv(cs)
                if b.idconv is not None:
                    add_conv(b.idconv)
        add_bn(spec.post_bn, spec.enc_out_c)
        for cs in spec.middle:
            add_conv(cs)
        for u in spec.unet:
            add_conv(u.shuf)
            add_bn(u.bn_prefix, u.x_in_c)
            add_conv(u.conv1)
            add_conv(u.conv2)
            if u.sa is not None:
                add(u.sa.gamma, (1,), True)     # a plain Parameter: fastai's wd_bn_bias=False does not exempt it
                for cs in u.sa.convs():
                    add_conv(cs)
        add_conv(spec.final_shuf)
        for cs in spec.final_res:
            add_conv(cs)
        add_conv(spec.head)
        self.total = (self._off + 3) // 4 * 4

    def names(self) -> List[str]:
        return [e.name for e in self.entries]

    def n_params(self) -> int:
        return sum(e.numel for e in self.entries)


def shuffle_row_of_co(nf4: int) -> List[int]:
    """GEMM row of torch output channel co for a PixelShuffle 1x1 conv with nf4 = 4*c outputs:
    torch channel 4*c_ + 2*i + j  ->  row (2*i + j) * c + c_   (so that each (i,j) phase is a contiguous channel run)."""
    c = nf4 // 4
    return [((co % 4) * c + co // 4) for co in range(nf4)]


def conv_flops(spec: NetSpec, size: int) -> int:
    """forward conv FLOPs per tile (2 FLOP/MAC, unpadded channels) — the numerator of SURVEY.md 8(d).  Odd extents halve
    upwards (3x3/s2/p1 convs, MaxPool(3,2,1), AvgPool(ceil_mode)); the decoder follows the skip sizes (crop)."""
    up = lambda v, st: (v + st - 1) // st
    total = 0
    s = up(size, 2)
    for i, cs in enumerate(spec.stem):
        total += 2 * s * s * cs.ni * cs.nf * 9
    skips = {2: s}
    s = up(s, 2)
    for si, st in enumerate(spec.stages):
        for b in st:
            so = up(s, b.stride)
            cur = s
            for cs in b.convpath:
                cur = up(cur, cs.stride)
                total += 2 * cur * cur * cs.ni * cs.nf * cs.ks * cs.ks
            if b.idconv is not None:
                total += 2 * so * so * b.idconv.ni * b.idconv.nf
            s = so
        skips[4 + si] = s
    for cs in spec.middle:
        total += 2 * s * s * cs.ni * cs.nf * 9
    for u in spec.unet:
        total += 2 * s * s * u.shuf.ni * u.shuf.nf
        s = skips[u.skip_child]
        total += 2 * s * s * (u.conv1.ni * u.conv1.nf + u.conv2.ni * u.conv2.nf) * 9
        if u.sa is not None:      # the three 1x1 convolutions only (the two batched attention products are not convs)
            total += 2 * s * s * sum(c.ni * c.nf for c in u.sa.convs())
    total += 2 * s * s * spec.final_shuf.ni * spec.final_shuf.nf
    s = size
    for cs in spec.final_res:
        total += 2 * s * s * cs.ni * cs.nf * 9
    total += 2 * s * s * spec.head.ni * spec.head.nf
    return total
