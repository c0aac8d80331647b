"""GPU diagnostic: one weight-gradient launch (for ncu). usage: one_wgrad.py CASE [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops
CASES = {"w100": (64, 256, 256, 100, 100, 3, True), "w256": (64, 64, 64, 256, 256, 3, True), "w96": (64, 128, 128, 96, 96, 3, True),
         "w32": (64, 128, 128, 32, 32, 3, False), "w1x1": (64, 128, 128, 96, 384, 1, True)}
case = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, H, W, Cin, Cout, ks, bias = CASES[case]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((N, H, W, ops.padc(Cin)), generator=g, device="cuda").to(torch.bfloat16)
dy = torch.randn((N, H, W, ops.padc(Cout)), generator=g, device="cuda").to(torch.bfloat16)
x[..., Cin:] = 0; dy[..., Cout:] = 0
dw = torch.zeros((Cout, Cin, ks, ks), device="cuda"); db = torch.zeros(Cout, device="cuda") if bias else None
taps = ops.taps_conv(ks)
plan = ops.WgradPlan(ops.view_nhwc(dy, Cout), [ops.view_nhwc(x, Cin)], taps, Cout, Cin, ks * ks, [t[3] for t in taps], dw.view(-1), db)
for _ in range(2): plan.run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps): plan.run()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / reps
print(f"{case}: {ms:.3f} ms (gemm+reduce) {plan.flops / ms / 1e9:.1f} TFLOP/s info={[(n, getattr(plan.info, n)) for n, _ in plan.info._fields_]}")
