"""Host-side operator layer over the C-ABI: views, tap tables and plan objects for the implicit-GEMM kernels.

Activations are torch bf16 tensors [N, H, W, Cp] (NHWC, channel pitch Cp a multiple of 8); torch only owns memory.
All arithmetic happens in libb2u.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, ConvInfo, View, WgradDesc, WgradInfo


# bench.py sets this to a list to time every GEMM launch with CUDA events on the launching stream (eager mode only):
# entries are (kind, plan, start_event, end_event).
PROFILE = None


def _timed(kind, plan, fn):
    if PROFILE is None:
        fn()
        return
    st = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    fn()
    b.record(st)
    PROFILE.append((kind, plan, a, b))


import os as _os
_PITCH16 = _os.environ.get("B2U_PITCH16") is not None    # A/B switch: plain 16-lane pitch everywhere


def padc(c: int) -> int:
    """channel pitch of NHWC activations / GEMM weights: whole 32-byte sectors (16 bf16).  TMA throughput halves when
    the innermost extent ends inside a sector (measured: Cin=104 -> 3.1 ms, Cin=112 -> 1.25 ms for the same launch)."""
    c16 = (c + 15) // 16 * 16
    # channel counts just above a multiple of 64 (100, 96, 99, 288, 292 ...) are pitched to the next multiple of 64 when
    # that costs <= 1/3 more lanes: the GEMM K loop then runs on whole 64-channel chunks (no partial TMA boxes, and the
    # 3x3 convolutions qualify for row-mode weight stages) - measured 0.91 ms vs 1.12 ms for the 100->100 3x3 launch
    c64 = (c + 63) // 64 * 64
    if c > 64 and c64 != c16 and (c64 - c) * 3 <= c and not _PITCH16:
        return c64
    return c16


def pad32(c: int) -> int:
    return (c + 31) // 32 * 32


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def view_nhwc(t: torch.Tensor, C_: Optional[int] = None, c_off: int = 0,
              parity: Optional[Tuple[int, int]] = None) -> View:
    """A b2u_view over a contiguous NHWC bf16 tensor: a channel slice [c_off, c_off+C) and optionally one of the four
    stride-2 parity planes (py, px): element (n, u, v, c) = t[n, 2u+py, 2v+px, c_off+c]."""
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    N, H, W, Cp = t.shape
    assert Cp % 8 == 0 and c_off % 8 == 0
    C_ = Cp - c_off if C_ is None else C_
    assert 0 < C_ <= Cp - c_off
    # the library rounds the channel extent of a view up to whole 64-lane chunks when the pitch allows (pad lanes are
    # zeros); a slice that does not start at lane 0 must therefore end on such a boundary itself
    assert c_off == 0 or C_ % 64 == 0 or c_off + (C_ + 63) // 64 * 64 <= Cp, "unsupported channel slice"
    v = View()
    if parity is None:
        v.ptr = t.data_ptr() + 2 * c_off
        v.C, v.W, v.H, v.N = C_, W, H, N
        v.sW, v.sH, v.sN = Cp, W * Cp, H * W * Cp
    else:
        py, px = parity
        v.ptr = t.data_ptr() + 2 * ((py * W + px) * Cp + c_off)
        v.C, v.W, v.H, v.N = C_, (W - px + 1) // 2, (H - py + 1) // 2, N
        v.sW, v.sH, v.sN = 2 * Cp, 2 * W * Cp, H * W * Cp
    return v


def null_view() -> View:
    return View()


Tap = Tuple[int, int, int, int]  # (view index, dy, dx, weight tap)


def taps_conv(ks: int) -> List[Tap]:
    """stride-1 'same' convolution: taps (r,s) read the single input view at offset (r-p, s-p)."""
    p = (ks - 1) // 2
    return [(0, r - p, s - p, r * ks + s) for r in range(ks) for s in range(ks)]


_S2 = {0: (-1, 1), 1: (0, 0), 2: (0, 1)}  # filter row r -> (plane offset, plane parity) for stride 2, pad 1


def taps_conv3_s2() -> List[Tap]:
    """3x3 / stride 2 / pad 1 over the four parity planes (view index = 2*py+px) of the input."""
    out = []
    for r in range(3):
        dy, py = _S2[r]
        for s in range(3):
            dx, px = _S2[s]
            out.append((2 * py + px, dy, dx, r * 3 + s))
    return out


def taps_avgpool_1x1() -> List[Tap]:
    """AvgPool2d(2) followed by a 1x1 conv == four taps (one per parity plane) sharing the 1x1 filter scaled by 1/4."""
    return [(2 * py + px, 0, 0, 0) for py in range(2) for px in range(2)]


def taps_dgrad_s2(py: int, px: int) -> List[Tap]:
    """Input-gradient of the 3x3/stride-2/pad-1 conv for the (py,px) parity plane of dX: a stride-1 conv over dY using
    the filter rows/cols of matching parity. Weight tap indices address the *flipped* dgrad layout (8 - t)."""
    rows = [(1, 0)] if py == 0 else [(0, 1), (2, 0)]
    cols = [(1, 0)] if px == 0 else [(0, 1), (2, 0)]
    return [(0, di, dj, 8 - (r * 3 + s)) for (r, di) in rows for (s, dj) in cols]


def _fill_taps(desc, taps: Sequence[Tap], with_w: bool = True):
    desc.num_taps = len(taps)
    for i, (a, dy, dx, w) in enumerate(taps):
        desc.tap_a[i], desc.tap_dy[i], desc.tap_dx[i] = a, dy, dx
        if with_w:
            desc.tap_w[i] = w


class ConvPlan:
    """One launch of the implicit-GEMM kernel with its TMA descriptors pre-encoded (b2u_conv_plan)."""

    def __init__(self, a_views: Sequence[View], out_view: View, w: torch.Tensor, w_cin: int, taps: Sequence[Tap],
                 scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
                 res: Optional[View] = None, res_mask: Optional[View] = None, zmask: Optional[View] = None,
                 relu: bool = False, stats: bool = False, out_f32: Optional[torch.Tensor] = None,
                 stats_ld: Optional[int] = None, fin: Optional[dict] = None, outs: Optional[Sequence[View]] = None,
                 w_batch_rows: int = 0, head: Optional[dict] = None):
        """fin (with stats=True): dict(count, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, scale,
        shift) of fp32 tensors -> the kernel's last CTA finalizes the BatchNorm statistics itself when the shape allows
        (self.fused_finalize tells the caller whether a separate b2u_bn_finalize launch is still needed)."""
        assert w.dtype == torch.bfloat16 and w.dim() == 3 and w.is_contiguous()
        lib = _lib.load()
        d = ConvDesc()
        for i, v in enumerate(a_views):
            d.a[i] = v
        d.num_a = len(a_views)
        d.out = out_view
        if outs:
            # N tile i (Cout / len(outs) channels) is stored to outs[i]; out_view gives the geometry and the total Cout
            assert 2 <= len(outs) <= 4
            d.num_out = len(outs)
            for i, v in enumerate(outs):
                d.out_nt[i] = v
        d.w = w.data_ptr()
        d.w_rows, d.w_taps, d.w_cinp = w.shape
        if w_batch_rows:
            # batched weights (torch.bmm on the implicit-GEMM kernel): w holds N blocks of w_batch_rows rows, image n of the
            # GEMM space multiplies block n; the first out.C rows of a block are used
            assert w.shape[0] == out_view.N * w_batch_rows and out_view.C <= w_batch_rows, (w.shape, out_view.N, w_batch_rows)
            d.w_rows = out_view.C
            d.w_batch_rows = w_batch_rows
        d.w_cin = w_cin
        _fill_taps(d, taps)
        for t in (scale, shift):
            if t is not None:
                assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= pad32(out_view.C), \
                    "scale/shift must hold pad32(Cout) floats (the epilogue reads whole 32-channel groups)"
        self._keep = [w, scale, shift, out_f32]
        d.scale = scale.data_ptr() if scale is not None else None
        d.shift = shift.data_ptr() if shift is not None else None
        if res is not None:
            d.res = res
        if res_mask is not None:
            d.res_mask = res_mask
        if zmask is not None:
            d.zmask = zmask
        flags = 0
        if relu:
            flags |= _lib.EPI_RELU
        if out_f32 is not None:
            assert out_f32.dtype == torch.float32 and out_f32.is_contiguous()
            flags |= _lib.EPI_OUT_F32
            d.out_f32 = out_f32.data_ptr()
            d.out_f32_ld = out_f32.shape[-1]
        if head is not None:
            # fused 1x1 head: dict(w = staged bf16 head weights [n_out, 1, ld], b = fp32 bias or None, out = fp32 logits
            # [N, H, W, ld_out], only = skip the bf16 output of this convolution)
            hw, hb, ho = head["w"], head.get("b"), head["out"]
            assert hw.dtype == torch.bfloat16 and hw.is_contiguous() and ho.dtype == torch.float32 and ho.is_contiguous()
            flags |= _lib.EPI_HEAD | (_lib.EPI_HEAD_ONLY if head.get("only") else 0)
            d.head_w, d.head_n, d.head_ld = hw.data_ptr(), hw.shape[0], hw.shape[-1]
            d.head_b = hb.data_ptr() if hb is not None else None
            d.out_f32, d.out_f32_ld = ho.data_ptr(), ho.shape[-1]
            self._keep += [hw, hb, ho]
        d.flags = flags
        self.stats = None
        self.fused_finalize = False
        if stats:
            info = ConvInfo()
            d.flags = flags | _lib.EPI_STATS   # the partial-row count depends on the statistics mode: query WITH the flag
            ld = stats_ld or padc(out_view.C)
            d.stats_ld = ld
            _lib.check(lib.b2u_conv_query(C.byref(d), C.byref(info)), "b2u_conv_query")
            self.stats = torch.zeros((info.stats_rows, 2, ld), dtype=torch.float32, device=w.device)
            d.stats = self.stats.data_ptr()
            if fin is not None and info.fused_finalize:
                self.fused_finalize = True
                self._fin_counter = torch.zeros(1, dtype=torch.int32, device=w.device)
                d.fin.counter = self._fin_counter.data_ptr()
                d.fin.count = float(fin["count"])
                d.fin.eps, d.fin.momentum = float(fin["eps"]), float(fin["momentum"])
                for k in ("gamma", "beta", "running_mean", "running_var", "mean", "invstd", "scale", "shift"):
                    t = fin.get(k)
                    if t is not None:
                        assert t.dtype == torch.float32 and t.is_contiguous()
                        setattr(d.fin, k, t.data_ptr())
                        self._keep.append(t)
        self.desc = d
        h = C.c_void_p()
        _lib.check(lib.b2u_conv_plan_create(C.byref(d), C.byref(h)), "b2u_conv_plan_create")
        self.handle = h
        self.info = ConvInfo()
        _lib.check(lib.b2u_conv_plan_info(h, C.byref(self.info)), "b2u_conv_plan_info")
        self._lib = lib

    def run(self, stream: Optional[int] = None):
        _timed("conv", self, lambda: _lib.check(
            self._lib.b2u_conv_run(self.handle, C.c_void_p(stream if stream is not None else stream_ptr())),
            "b2u_conv_run"))

    @property
    def flops(self) -> int:
        """algorithmic FLOPs of this launch: 2 * pixels * Cout * Cin * taps (unpadded channels)"""
        d = self.desc
        return 2 * d.out.N * d.out.H * d.out.W * d.out.C * getattr(self, "alg_cin", d.w_cin) * d.num_taps

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            self._lib.b2u_conv_plan_destroy(h)
            self.handle = None


class WgradPlan:
    """Weight-gradient GEMM + deterministic split reduction into a torch-layout fp32 gradient."""

    def __init__(self, dy_view: View, a_views: Sequence[View], taps: Sequence[Tap], Cout: int, Cin: int,
                 ksize: int, tap_kidx: Sequence[int], dw: torch.Tensor, db: Optional[torch.Tensor] = None,
                 row_perm: Optional[torch.Tensor] = None, alpha: float = 1.0,
                 workspace: Optional[torch.Tensor] = None):
        lib = _lib.load()
        d = WgradDesc()
        d.dy = dy_view
        for i, v in enumerate(a_views):
            d.a[i] = v
        d.num_a = len(a_views)
        _fill_taps(d, taps, with_w=False)
        d.Cout, d.Cin = Cout, Cin
        d.want_bias = 1 if db is not None else 0
        info = WgradInfo()
        _lib.check(lib.b2u_wgrad_query(C.byref(d), C.byref(info)), "b2u_wgrad_query")
        dev = dw.device
        if workspace is None or workspace.numel() * 4 < info.partial_bytes:
            workspace = torch.empty((info.partial_bytes + 3) // 4, dtype=torch.float32, device=dev)
        self.workspace = workspace
        d.partial = workspace.data_ptr()
        d.partial_bytes = workspace.numel() * 4
        h = C.c_void_p()
        _lib.check(lib.b2u_wgrad_plan_create(C.byref(d), C.byref(h)), "b2u_wgrad_plan_create")
        self.handle, self.info, self.desc = h, info, d
        assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == Cout * Cin * ksize
        self.dw, self.db = dw, db
        self.kidx = torch.tensor(list(tap_kidx), dtype=torch.int32, device=dev)
        self.row_perm = row_perm
        self.alpha, self.ksize, self.Cout, self.Cin, self.ntaps = alpha, ksize, Cout, Cin, len(taps)
        self._lib = lib

    @staticmethod
    def query(dy_view: View, a_views: Sequence[View], taps: Sequence[Tap], Cout: int, Cin: int,
              want_bias: bool) -> WgradInfo:
        lib = _lib.load()
        d = WgradDesc()
        d.dy = dy_view
        for i, v in enumerate(a_views):
            d.a[i] = v
        d.num_a = len(a_views)
        _fill_taps(d, taps, with_w=False)
        d.Cout, d.Cin, d.want_bias = Cout, Cin, int(want_bias)
        info = WgradInfo()
        _lib.check(lib.b2u_wgrad_query(C.byref(d), C.byref(info)), "b2u_wgrad_query")
        return info

    @property
    def flops(self) -> int:
        d = self.desc
        return 2 * d.dy.N * d.dy.H * d.dy.W * d.Cout * d.Cin * d.num_taps

    def run(self, stream: Optional[int] = None):
        s = C.c_void_p(stream if stream is not None else stream_ptr())
        _timed("wgrad", self, lambda: _lib.check(self._lib.b2u_wgrad_run(self.handle, s), "b2u_wgrad_run"))
        _lib.check(self._lib.b2u_wgrad_reduce(
            self.workspace.data_ptr(), self.info.splits, self.ntaps, self.info.co_pad, self.info.ci_pad, self.Cout,
            self.Cin, self.ksize, self.kidx.data_ptr(),
            self.row_perm.data_ptr() if self.row_perm is not None else None, self.alpha, self.dw.data_ptr(),
            self.db.data_ptr() if self.db is not None else None, 1 if self.db is not None else 0, s),
            "b2u_wgrad_reduce")

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            self._lib.b2u_wgrad_plan_destroy(h)
            self.handle = None
