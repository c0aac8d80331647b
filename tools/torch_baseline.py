"""Secondary, informative bar: the oracle module under stock torch + cuDNN, bf16 autocast, channels_last,
on the same B200 — the strongest off-the-shelf implementation available on the box (NOT part of the product)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.unet_oracle import make_oracle, weighted_ce
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.backends.cudnn.benchmark = True
m = make_oracle("xresnet34", 4, 2).cuda().train().to(memory_format=torch.channels_last)
opt = torch.optim.SGD(m.parameters(), lr=1e-3)
x = torch.rand(B, 4, 256, 256, device="cuda").contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 2, (B, 256, 256), device="cuda")
w = torch.full((2,), 0.5, device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x)
    loss = weighted_ce(out.float(), y, w)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"torch+cuDNN bf16 autocast channels_last: batch {B}: {ms:.2f} ms/step = {B / ms * 1e3:.1f} tiles/s")
