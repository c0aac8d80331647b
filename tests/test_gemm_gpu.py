"""GPU parity of the tcgen05 implicit-GEMM kernels (fprop / dgrad / wgrad) against torch fp32 convolutions computed on
the same bf16-rounded operands.  Tolerance: max|a-b| / max|b| <= 1e-2 for bf16 outputs (north_star bf16 bar),
<= 2e-3 for fp32 outputs (wgrad, fp32 logits) — inputs are bf16 so products are exact and only accumulation order differs.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def to_nhwc(x_nchw, cp=None):
    from unet_b200.ops import padc
    n, c, h, w = x_nchw.shape
    cp = cp or padc(c)
    out = torch.zeros((n, h, w, cp), dtype=torch.bfloat16, device=x_nchw.device)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def gemm_weights(w):
    """torch [Cout,Cin,kh,kw] fp32 -> bf16 [Cout][kh*kw][CinP]"""
    from unet_b200.ops import padc
    co, ci, kh, kw = w.shape
    out = torch.zeros((co, kh * kw, padc(ci)), dtype=torch.bfloat16, device=w.device)
    out[..., :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).to(torch.bfloat16)
    return out


def dgrad_weights(w):
    """torch [Cout,Cin,kh,kw] -> bf16 [Cin][kk (flipped)][CoutP]"""
    from unet_b200.ops import padc
    co, ci, kh, kw = w.shape
    out = torch.zeros((ci, kh * kw, padc(co)), dtype=torch.bfloat16, device=w.device)
    out[..., :co] = w.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, kh * kw, co).to(torch.bfloat16)
    return out


def padvec(v):
    """per-channel epilogue vectors are read in whole 32-float groups"""
    from unet_b200.ops import pad32
    out = torch.zeros(pad32(v.numel()), dtype=torch.float32, device=v.device)
    out[:v.numel()] = v
    return out


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16).float()


CONV_CASES = [
    # N, Cin, Cout, H, W, ks
    (2, 64, 64, 16, 16, 3),
    (2, 64, 64, 64, 64, 3),
    (3, 128, 128, 32, 32, 3),
    (4, 256, 256, 16, 16, 3),
    (4, 512, 512, 8, 8, 3),
    (1, 100, 100, 128, 128, 3),
    (1, 192, 96, 64, 64, 3),
    (2, 32, 64, 128, 128, 3),
    (2, 512, 1024, 8, 8, 3),
    (2, 1024, 512, 8, 8, 3),
    (2, 384, 768, 32, 32, 1),
    (1, 96, 384, 64, 64, 1),
    (2, 4, 32, 64, 64, 3),
    (1, 64, 64, 25, 25, 3),     # ragged (the reference's default 400-px tiles reach 25x25)
    (3, 48, 40, 13, 50, 3),     # ragged, odd channel counts
]


@pytest.mark.parametrize("N,Cin,Cout,H,W,ks", CONV_CASES)
def test_conv_fprop_s1(N, Cin, Cout, H, W, ks):
    from unet_b200 import ops
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, ks, ks, seed=2, scale=(Cin * ks * ks) ** -0.5)
    b = rnd(Cout, seed=3)
    ref = F.relu(F.conv2d(x, w, b, padding=(ks - 1) // 2))
    xa = to_nhwc(x)
    ya = torch.full((N, H, W, ops.padc(Cout)), 7.0, dtype=torch.bfloat16, device="cuda")
    plan = ops.ConvPlan([ops.view_nhwc(xa, Cin)], ops.view_nhwc(ya, Cout), gemm_weights(w), Cin, ops.taps_conv(ks),
                        shift=padvec(b), relu=True)
    plan.run()
    torch.cuda.synchronize()
    got = ya[..., :Cout].permute(0, 3, 1, 2).float()
    e = rel_err(got, ref)
    assert e <= 1e-2, f"rel err {e} info={[(n, getattr(plan.info, n)) for n, _ in plan.info._fields_]}"
    if ops.padc(Cout) != Cout:  # pad lanes up to the 16-byte granule are written as zeros (TMA stores whole granules)
        assert (ya[..., Cout:] == 0.0).all()


def test_conv_fprop_stats_and_residual():
    from unet_b200 import ops
    N, Cin, Cout, H, W = 2, 64, 128, 32, 32
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=(Cin * 9) ** -0.5)
    r = rnd(N, Cout, H, W, seed=4)
    m = rnd(N, Cout, H, W, seed=5)
    z = rnd(N, Cout, H, W, seed=6)
    sc = rnd(Cout, seed=7).abs() + 0.5
    sh = rnd(Cout, seed=8)
    ref = F.conv2d(x, w, None, padding=1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    ref = ref + r * (m > 0)
    ref = F.relu(ref) * (z > 0)
    xa, ra, ma, za = to_nhwc(x), to_nhwc(r), to_nhwc(m), to_nhwc(z)
    ya = torch.zeros((N, H, W, Cout), dtype=torch.bfloat16, device="cuda")
    plan = ops.ConvPlan([ops.view_nhwc(xa)], ops.view_nhwc(ya), gemm_weights(w), Cin, ops.taps_conv(3),
                        scale=padvec(sc), shift=padvec(sh), res=ops.view_nhwc(ra),
                        res_mask=ops.view_nhwc(ma), zmask=ops.view_nhwc(za), relu=True, stats=True)
    plan.run()
    torch.cuda.synchronize()
    got = ya.permute(0, 3, 1, 2).float()
    assert rel_err(got, ref) <= 1e-2
    s = plan.stats.sum(0)  # [2, ld]
    assert rel_err(s[0, :Cout], got.sum((0, 2, 3))) <= 1e-4
    assert rel_err(s[1, :Cout], (got * got).sum((0, 2, 3))) <= 1e-4


@pytest.mark.parametrize("N,Cin,Cout,H,W", [(2, 64, 128, 64, 64), (2, 4, 32, 64, 64), (1, 256, 512, 16, 16)])
def test_conv_fprop_stride2(N, Cin, Cout, H, W):
    from unet_b200 import ops
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=(Cin * 9) ** -0.5)
    ref = F.conv2d(x, w, None, stride=2, padding=1)
    xa = to_nhwc(x)
    ya = torch.zeros((N, H // 2, W // 2, ops.padc(Cout)), dtype=torch.bfloat16, device="cuda")
    views = [ops.view_nhwc(xa, Cin, parity=(py, px)) for py in range(2) for px in range(2)]
    plan = ops.ConvPlan(views, ops.view_nhwc(ya, Cout), gemm_weights(w), Cin, ops.taps_conv3_s2())
    plan.run()
    torch.cuda.synchronize()
    assert rel_err(ya[..., :Cout].permute(0, 3, 1, 2), ref) <= 1e-2


def test_avgpool_1x1_fused():
    from unet_b200 import ops
    N, Cin, Cout, H, W = 2, 64, 128, 32, 32
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 1, 1, seed=2, scale=Cin ** -0.5)
    ref = F.conv2d(F.avg_pool2d(x, 2, ceil_mode=True), w)
    xa = to_nhwc(x)
    ya = torch.zeros((N, H // 2, W // 2, Cout), dtype=torch.bfloat16, device="cuda")
    views = [ops.view_nhwc(xa, Cin, parity=(py, px)) for py in range(2) for px in range(2)]
    plan = ops.ConvPlan(views, ops.view_nhwc(ya), gemm_weights(w * 0.25), Cin, ops.taps_avgpool_1x1())
    plan.run()
    torch.cuda.synchronize()
    assert rel_err(ya.permute(0, 3, 1, 2), ref) <= 1e-2


@pytest.mark.parametrize("N,Cin,Cout,H,W,ks", [(2, 64, 64, 32, 32, 3), (1, 100, 100, 64, 64, 3), (2, 384, 768, 16, 16, 1)])
def test_dgrad_s1(N, Cin, Cout, H, W, ks):
    from unet_b200 import ops
    w = rnd(Cout, Cin, ks, ks, seed=2, scale=(Cout * ks * ks) ** -0.5)
    dy = rnd(N, Cout, H, W, seed=3)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, padding=(ks - 1) // 2)
    dya = to_nhwc(dy)
    dxa = torch.zeros((N, H, W, ops.padc(Cin)), dtype=torch.bfloat16, device="cuda")
    plan = ops.ConvPlan([ops.view_nhwc(dya, Cout)], ops.view_nhwc(dxa, Cin), dgrad_weights(w), Cout, ops.taps_conv(ks))
    plan.run()
    torch.cuda.synchronize()
    assert rel_err(dxa[..., :Cin].permute(0, 3, 1, 2), ref) <= 1e-2


def test_dgrad_s2():
    from unet_b200 import ops
    N, Cin, Cout, H, W = 2, 64, 128, 32, 32
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=(Cout * 9) ** -0.5)
    dy = rnd(N, Cout, H // 2, W // 2, seed=3)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, stride=2, padding=1)
    dya = to_nhwc(dy)
    dxa = torch.zeros((N, H, W, Cin), dtype=torch.bfloat16, device="cuda")
    wd = dgrad_weights(w)
    for py in range(2):
        for px in range(2):
            plan = ops.ConvPlan([ops.view_nhwc(dya)], ops.view_nhwc(dxa, parity=(py, px)), wd, Cout,
                                ops.taps_dgrad_s2(py, px))
            plan.run()
    torch.cuda.synchronize()
    assert rel_err(dxa.permute(0, 3, 1, 2), ref) <= 1e-2


def test_conv_out_f32_head():
    from unet_b200 import ops
    N, Cin, Cout, H, W = 2, 100, 2, 64, 64
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 1, 1, seed=2, scale=Cin ** -0.5)
    b = rnd(Cout, seed=3)
    ref = F.conv2d(x, w, b)
    xa = to_nhwc(x)
    out = torch.zeros((N, H, W, 8), dtype=torch.float32, device="cuda")
    ov = ops.view_nhwc(torch.zeros((N, H, W, 8), dtype=torch.bfloat16, device="cuda"), Cout)
    plan = ops.ConvPlan([ops.view_nhwc(xa, Cin)], ov, gemm_weights(w), Cin, ops.taps_conv(1), shift=padvec(b),
                        out_f32=out)
    plan.run()
    torch.cuda.synchronize()
    assert rel_err(out[..., :Cout].permute(0, 3, 1, 2), ref) <= 2e-3


WGRAD_CASES = [
    # N, Cin, Cout, H, W, ks, bias
    (2, 64, 64, 32, 32, 3, True),
    (2, 100, 100, 64, 64, 3, True),
    (4, 512, 512, 8, 8, 3, True),
    (2, 128, 256, 16, 16, 3, False),
    (2, 384, 768, 16, 16, 1, True),
    (1, 100, 2, 64, 64, 1, True),
    (2, 4, 32, 32, 32, 3, False),
    (1, 48, 40, 13, 50, 3, True),
]


@pytest.mark.parametrize("N,Cin,Cout,H,W,ks,bias", WGRAD_CASES)
def test_wgrad_s1(N, Cin, Cout, H, W, ks, bias):
    from unet_b200 import ops
    x = rnd(N, Cin, H, W, seed=1)
    dy = rnd(N, Cout, H, W, seed=3)
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, ks, ks), dy, padding=(ks - 1) // 2)
    xa, dya = to_nhwc(x), to_nhwc(dy)
    dw = torch.zeros((Cout, Cin, ks, ks), dtype=torch.float32, device="cuda")
    db = torch.zeros(Cout, dtype=torch.float32, device="cuda") if bias else None
    taps = ops.taps_conv(ks)
    plan = ops.WgradPlan(ops.view_nhwc(dya, Cout), [ops.view_nhwc(xa, Cin)], taps, Cout, Cin, ks * ks,
                         [t[3] for t in taps], dw, db)
    plan.run()
    torch.cuda.synchronize()
    e = rel_err(dw, ref)
    assert e <= 2e-3, f"rel err {e} info={[(n, getattr(plan.info, n)) for n, _ in plan.info._fields_]}"
    if bias:
        assert rel_err(db, dy.sum((0, 2, 3))) <= 2e-3


def test_wgrad_s2():
    from unet_b200 import ops
    N, Cin, Cout, H, W = 2, 64, 128, 32, 32
    x = rnd(N, Cin, H, W, seed=1)
    dy = rnd(N, Cout, H // 2, W // 2, seed=3)
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, 3, 3), dy, stride=2, padding=1)
    xa, dya = to_nhwc(x), to_nhwc(dy)
    dw = torch.zeros((Cout, Cin, 3, 3), dtype=torch.float32, device="cuda")
    taps = ops.taps_conv3_s2()
    views = [ops.view_nhwc(xa, Cin, parity=(py, px)) for py in range(2) for px in range(2)]
    plan = ops.WgradPlan(ops.view_nhwc(dya), views, taps, Cout, Cin, 9, [t[3] for t in taps], dw)
    plan.run()
    torch.cuda.synchronize()
    assert rel_err(dw, ref) <= 2e-3


@pytest.mark.parametrize("N,Cin,Cout,H,W,stride", [(8, 64, 128, 64, 64, 1), (6, 32, 64, 128, 128, 1), (8, 64, 128, 64, 64, 2),
                                                    (4, 256, 512, 32, 32, 2)])
def test_conv_stats_many_tiles_per_cta(N, Cin, Cout, H, W, stride):
    """BN partial sums accumulated on chip over all tiles of a persistent CTA (more tiles than SMs), one and two N tiles."""
    from unet_b200 import ops
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=(Cin * 9) ** -0.5)
    ref = F.conv2d(x, w, None, stride=stride, padding=1)
    xa = to_nhwc(x)
    ya = torch.zeros((N, H // stride, W // stride, ops.padc(Cout)), dtype=torch.bfloat16, device="cuda")
    if stride == 1:
        views, taps = [ops.view_nhwc(xa, Cin)], ops.taps_conv(3)
    else:
        views, taps = [ops.view_nhwc(xa, Cin, parity=(py, px)) for py in range(2) for px in range(2)], ops.taps_conv3_s2()
    plan = ops.ConvPlan(views, ops.view_nhwc(ya, Cout), gemm_weights(w), Cin, taps, stats=True)
    plan.run()
    plan.run()     # a second launch must not accumulate onto the first
    torch.cuda.synchronize()
    got = ya[..., :Cout].permute(0, 3, 1, 2).float()
    assert rel_err(got, ref) <= 1e-2
    s = plan.stats.double().sum(0)
    assert rel_err(s[0, :Cout].float(), got.double().sum((0, 2, 3)).float()) <= 1e-4, plan.stats.shape
    assert rel_err(s[1, :Cout].float(), (got.double() ** 2).sum((0, 2, 3)).float()) <= 1e-4


@pytest.mark.parametrize("N,Cin,Cout,H,W", [(8, 64, 64, 64, 64), (4, 256, 512, 8, 8), (2, 32, 100, 40, 24)])
def test_conv_fused_bn_finalize(N, Cin, Cout, H, W):
    """The conv kernel's last CTA finalizes the BatchNorm batch statistics (b2u_bn_fin): mean / invstd / scale / shift and
    the running statistics must equal torch's training-mode batch_norm on the stored (bf16) conv output."""
    from unet_b200 import ops
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=(Cin * 9) ** -0.5)
    xa = to_nhwc(x)
    ya = torch.zeros((N, H, W, ops.padc(Cout)), dtype=torch.bfloat16, device="cuda")
    f = lambda fill=0.0: torch.full((ops.pad32(Cout),), fill, dtype=torch.float32, device="cuda")
    gamma, beta = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda") * 0.1
    rm0, rv0 = torch.randn(Cout, device="cuda") * 0.1, torch.rand(Cout, device="cuda") + 0.5
    rm, rv = rm0.clone(), rv0.clone()
    mean, invstd, scale, shift = f(), f(), f(), f()
    plan = ops.ConvPlan([ops.view_nhwc(xa, Cin)], ops.view_nhwc(ya, Cout), gemm_weights(w), Cin, ops.taps_conv(3), stats=True,
                        fin=dict(count=N * H * W, gamma=gamma, beta=beta, eps=1e-5, momentum=0.1, running_mean=rm,
                                 running_var=rv, mean=mean, invstd=invstd, scale=scale, shift=shift))
    assert plan.fused_finalize
    plan.run()
    torch.cuda.synchronize()
    got = ya[..., :Cout].permute(0, 3, 1, 2).float()
    rm_ref, rv_ref = rm0.clone(), rv0.clone()
    y_ref = F.batch_norm(got, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    m = got.double().mean((0, 2, 3))
    v = got.double().var((0, 2, 3), unbiased=False)
    assert rel_err(mean[:Cout], m.float()) <= 1e-4
    assert rel_err(invstd[:Cout], (1.0 / torch.sqrt(v + 1e-5)).float()) <= 1e-4
    assert rel_err(rm, rm_ref) <= 1e-4 and rel_err(rv, rv_ref) <= 1e-4
    y = got * scale[:Cout].view(1, -1, 1, 1) + shift[:Cout].view(1, -1, 1, 1)
    assert rel_err(y, y_ref) <= 1e-4
    # a second launch restarts from a clean counter and applies the momentum update once more
    plan.run()
    torch.cuda.synchronize()
    F.batch_norm(got, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    assert rel_err(rm, rm_ref) <= 1e-4 and rel_err(rv, rv_ref) <= 1e-4
    assert plan._fin_counter.item() == 0


@pytest.mark.parametrize("N,Cin,cu,H,W", [(2, 96, 96, 32, 32), (1, 64, 64, 16, 24), (2, 96, 96, 7, 9)])
def test_conv_pixel_shuffle_store(N, Cin, cu, H, W):
    """num_out = 4: the 1x1 PixelShuffle_ICNR convolution (+bias, ReLU) stores its four (i,j) phases straight into the
    stride-2 parity planes of the upsampled tensor == F.pixel_shuffle(relu(conv(x))) (fastai layers.py
    PixelShuffle_ICNR, blur=False); lanes past cu are zero-filled by the store; b2u_copy_lanes then appends the image
    bands (MergeLayer dense) and b2u_shuffle_bwd_from_cat reproduces the plain shuffle backward bit for bit."""
    from unet_b200 import _lib, ops
    from unet_b200.layout import shuffle_row_of_co
    L = _lib.load()
    Cout = 4 * cu
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 1, 1, seed=2, scale=Cin ** -0.5)
    b = rnd(Cout, seed=3)
    ref = F.pixel_shuffle(F.relu(F.conv2d(x, w, b)), 2)                       # [N, cu, 2H, 2W]
    roc = shuffle_row_of_co(Cout)                                             # torch channel co -> GEMM row (i,j,c)
    wg = torch.zeros((Cout, 1, ops.padc(Cin)), dtype=torch.bfloat16, device="cuda")
    bg = torch.zeros(ops.pad32(Cout), dtype=torch.float32, device="cuda")
    for co, r in enumerate(roc):
        wg[r, 0, :Cin] = w[co, :, 0, 0].to(torch.bfloat16)
        bg[r] = b[co]
    xa = to_nhwc(x)
    ldc = ops.padc(cu + 4)
    cat = torch.full((N, 2 * H, 2 * W, ldc), 7.0, dtype=torch.bfloat16, device="cuda")     # poisoned: every lane is written
    outs = [ops.view_nhwc(cat, cu, parity=(i, j)) for i in range(2) for j in range(2)]
    geom = ops.view_nhwc(cat, cu, parity=(0, 0))
    geom.C = Cout
    plan = ops.ConvPlan([ops.view_nhwc(xa, Cin)], geom, wg, Cin, ops.taps_conv(1), shift=bg, relu=True, outs=outs)
    plan.run()
    img = rnd(N, 4, 2 * H, 2 * W, seed=4)
    xi = to_nhwc(img)
    lanes = min(xi.shape[-1], ldc - cu)
    _lib.check(L.b2u_copy_lanes(xi.data_ptr(), xi.shape[-1], 0, cat.data_ptr(), ldc, cu, lanes, N * 4 * H * W,
                                ops.stream_ptr()), "b2u_copy_lanes")
    torch.cuda.synchronize()
    got = cat[..., :cu].permute(0, 3, 1, 2).float()
    assert rel_err(got, ref) <= 1e-2
    assert torch.equal(cat[..., cu:cu + 4], xi[..., :4]) and not cat[..., cu + 4:cu + lanes].any()
    # backward of the shuffle with the mask taken from cat == the plain kernel fed with the pre-shuffle activation
    P = torch.zeros((N, H, W, ops.padc(Cout)), dtype=torch.bfloat16, device="cuda")
    for ij in range(4):
        P[..., ij * cu:(ij + 1) * cu] = cat[:, ij // 2::2, ij % 2::2, :cu]
    dcat = to_nhwc(rnd(N, cu + 4, 2 * H, 2 * W, seed=5), ldc)
    d1, d2 = torch.zeros_like(P), torch.zeros_like(P)
    _lib.check(L.b2u_shuffle_bwd(dcat.data_ptr(), ldc, P.data_ptr(), d1.data_ptr(), P.shape[-1], cu, 0, N, H, W,
                                 ops.stream_ptr()), "b2u_shuffle_bwd")
    _lib.check(L.b2u_shuffle_bwd_from_cat(dcat.data_ptr(), cat.data_ptr(), ldc, d2.data_ptr(), P.shape[-1], cu, N, H, W,
                                          ops.stream_ptr()), "b2u_shuffle_bwd_from_cat")
    torch.cuda.synchronize()
    assert torch.equal(d1, d2)


@pytest.mark.parametrize("N,h,w,Cin,rows", [(3, 16, 16, 48, 256), (2, 32, 32, 1024, 384), (4, 8, 8, 64, 48),
                                             (2, 25, 25, 48, 625), (3, 16, 16, 256, 48)])
def test_conv_batched_weights_is_bmm(N, h, w, Cin, rows):
    """b2u_conv_desc.w_batch_rows: out[img] = A[img] @ W[img]^T - torch.bmm on the implicit-GEMM kernel (the batched
    products of fastai's SelfAttention, forward and backward shapes; one image smaller than a 128-pixel tile)."""
    from unet_b200 import ops
    n = h * w
    a = rnd(N, n, Cin, seed=5, scale=Cin ** -0.5)
    wt = rnd(N, rows, Cin, seed=6)
    ref = torch.bmm(a, wt.transpose(1, 2))                                   # [N, n, rows]
    ld_a, ld_o = ops.padc(Cin), ops.padc(rows)
    at = torch.zeros((N, h, w, ld_a), dtype=torch.bfloat16, device="cuda")
    at[..., :Cin] = a.view(N, h, w, Cin).to(torch.bfloat16)
    wb = torch.zeros((N * rows, 1, ld_a), dtype=torch.bfloat16, device="cuda")
    wb[:, 0, :Cin] = wt.reshape(N * rows, Cin).to(torch.bfloat16)
    out = torch.full((N, h, w, ld_o), 3.0, dtype=torch.bfloat16, device="cuda")
    plan = ops.ConvPlan([ops.view_nhwc(at, Cin)], ops.view_nhwc(out, rows), wb, Cin, ops.taps_conv(1), w_batch_rows=rows)
    assert plan.info.tile_n == 1
    plan.run()
    torch.cuda.synchronize()
    e = rel_err(out.view(N, n, ld_o)[..., :rows], ref)
    assert e <= 1e-2, (e, [(k, getattr(plan.info, k)) for k, _ in plan.info._fields_])


@pytest.mark.parametrize("N,C,H,W,n_out,only,res", [(2, 100, 64, 64, 2, False, True), (1, 99, 32, 48, 2, True, True),
                                                     (2, 64, 16, 16, 8, False, False), (1, 128, 32, 32, 1, True, False)])
def test_conv_fused_head(N, C, H, W, n_out, only, res):
    """B2U_EPI_HEAD: the 1x1 head (layers.12) in the epilogue of the preceding 3x3 convolution: logits == head(bf16(out));
    with B2U_EPI_HEAD_ONLY the bf16 output is not written at all."""
    from unet_b200 import ops
    x = rnd(N, C, H, W, seed=1)
    w = rnd(C, C, 3, 3, seed=2, scale=(C * 9) ** -0.5)
    b = rnd(C, seed=3)
    hw_ = rnd(n_out, C, 1, 1, seed=4, scale=C ** -0.5)
    hb = rnd(n_out, seed=5)
    y_ref = F.conv2d(x, w, b, padding=1)
    if res:
        y_ref = y_ref + x
    y_ref = F.relu(y_ref).to(torch.bfloat16).float()
    logits_ref = F.conv2d(y_ref, hw_, hb)
    xa = to_nhwc(x)
    ya = torch.full((N, H, W, ops.padc(C)), 7.0, dtype=torch.bfloat16, device="cuda")
    lg = torch.full((N, H, W, n_out), 9.0, dtype=torch.float32, device="cuda")
    plan = ops.ConvPlan([ops.view_nhwc(xa, C)], ops.view_nhwc(ya, C), gemm_weights(w), C, ops.taps_conv(3),
                        shift=padvec(b), relu=True, res=ops.view_nhwc(xa, C) if res else None,
                        head=dict(w=gemm_weights(hw_), b=padvec(hb), out=lg, only=only))
    plan.run()
    torch.cuda.synchronize()
    assert rel_err(lg.permute(0, 3, 1, 2), logits_ref) <= 5e-3      # (a rounding-boundary flip of one bf16 output moves a logit by ~1e-3)
    if only:
        assert (ya == 7.0).all()
    else:
        assert rel_err(ya[..., :C].permute(0, 3, 1, 2), y_ref) <= 1e-2
