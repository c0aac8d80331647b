"""GPU diagnostic: run single implicit-GEMM launches (for ncu). usage: one_conv.py CASE [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops

CASES = {
    # name: (N, H, W, Cin, Cout, ks, relu, zmask, res, bias)
    "head_dgrad": (64, 256, 256, 2, 100, 1, False, True, False, False),
    "res100": (64, 256, 256, 100, 100, 3, True, False, True, True),
    "c256": (64, 64, 64, 256, 256, 3, True, False, False, True),
    "shuf96": (64, 128, 128, 96, 384, 1, True, False, False, True),
    "c64": (64, 64, 64, 64, 64, 3, False, False, False, False),
    "c128_100": (64, 256, 256, 128, 100, 3, True, False, False, True),
    "c128_112": (64, 256, 256, 128, 112, 3, True, False, False, True),
    "c128_128": (64, 256, 256, 128, 128, 3, True, False, False, True),
    "c64_100": (64, 256, 256, 64, 100, 3, True, False, False, True),
    "c100_100": (64, 256, 256, 100, 100, 3, True, False, False, True),
    "c112_112": (64, 256, 256, 112, 112, 3, True, False, False, True),
    "c96_96": (64, 128, 128, 96, 96, 3, True, False, False, True),
    "c128_128s": (64, 128, 128, 128, 128, 3, True, False, False, True),
    # encoder shapes; suffix s = BN statistics in the epilogue, f = statistics + fused finalize (11th field: 0 / 1 / 2)
    "e64": (64, 64, 64, 64, 64, 3, False, False, False, False, 0),
    "e64s": (64, 64, 64, 64, 64, 3, False, False, False, False, 1),
    "e64f": (64, 64, 64, 64, 64, 3, False, False, False, False, 2),
    "e128": (64, 32, 32, 128, 128, 3, False, False, False, False, 0),
    "e128f": (64, 32, 32, 128, 128, 3, False, False, False, False, 2),
    "e256": (64, 16, 16, 256, 256, 3, False, False, False, False, 0),
    "e256s": (64, 16, 16, 256, 256, 3, False, False, False, False, 1),
    "e256f": (64, 16, 16, 256, 256, 3, False, False, False, False, 2),
    "e512": (64, 8, 8, 512, 512, 3, False, False, False, False, 0),
    "e512f": (64, 8, 8, 512, 512, 3, False, False, False, False, 2),
    "s32": (64, 128, 128, 32, 32, 3, False, False, False, False, 0),
    "s32f": (64, 128, 128, 32, 32, 3, False, False, False, False, 2),
    "s32_64": (64, 128, 128, 32, 64, 3, False, False, False, False, 0),
    "s32_64f": (64, 128, 128, 32, 64, 3, False, False, False, False, 2),
    "s64_32": (64, 128, 128, 64, 32, 3, False, False, False, False, 0),
}
case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pitch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N, H, W, Cin, Cout, ks, relu, zm, res, bias = CASES[case][:10]
st = CASES[case][10] if len(CASES[case]) > 10 else 0
g = torch.Generator(device="cuda").manual_seed(0)
pi = max(pitch, ops.padc(Cin)); po = max(pitch, ops.padc(Cout))
x = torch.randn((N, H, W, pi), generator=g, device="cuda").to(torch.bfloat16)
w = (torch.randn((Cout, ks * ks, ops.padc(Cin)), generator=g, device="cuda") * (Cin * ks * ks) ** -0.5).to(torch.bfloat16)
y = torch.zeros((N, H, W, po), dtype=torch.bfloat16, device="cuda")
z = torch.randn((N, H, W, po), generator=g, device="cuda").to(torch.bfloat16)
b = torch.randn(ops.pad32(Cout), generator=g, device="cuda")
plan = ops.ConvPlan([ops.view_nhwc(x, Cin)], ops.view_nhwc(y, Cout), w, Cin, ops.taps_conv(ks),
                    shift=b if bias else None, relu=relu, zmask=ops.view_nhwc(z, Cout) if zm else None,
                    res=ops.view_nhwc(z, Cout) if res else None, stats=st > 0,
                    fin=dict(count=N * H * W, eps=1e-5, momentum=0.1,
                             **{k: torch.ones(po, device="cuda") for k in ("gamma", "beta", "running_mean", "running_var", "mean",
                                                                           "invstd", "scale", "shift")}) if st == 2 else None)
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    plan.run()
e.record()
torch.cuda.synchronize()
ms = a.elapsed_time(e) / reps
print(f"{case} pitch {pi}/{po}: {ms:.3f} ms  {plan.flops / ms / 1e9:.1f} TFLOP/s  info={[(n, getattr(plan.info, n)) for n, _ in plan.info._fields_]}")
