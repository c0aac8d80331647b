"""GPU diagnostic: stored activation gradients of the plan vs the teacher-forced emulation, tensor by tensor in backward
order - locates the first op whose backward differs.  usage: grad_bisect.py arch n_in n_out size batch"""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle.unet_oracle import make_oracle, weighted_ce
from oracle.bf16_emulation import emulated_forward
from unet_b200.network import UNetB200
from unet_b200.synth import aerial_like_tiles
from parity_util import plan_taps, rel, cosine
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
a = sys.argv
arch, n_in, n_out, size, batch = a[1], int(a[2]), int(a[3]), int(a[4]), int(a[5])
oracle = make_oracle(arch, n_in, n_out).cuda().train()
net = UNetB200(arch, n_in, n_out, (size, size), batch, training=True)
net.load_state_dict(oracle.state_dict())
x_u8, y = aerial_like_tiles(batch, n_in, size, size, n_out)
x_u8, y = x_u8.cuda(), y.cuda()
x, yl = x_u8.float() / 255, y.long()
w = torch.full((n_out,), 1.0 / n_out, device="cuda")
net.set_input(x_u8); net.set_labels(y); net.forward(); net.loss_and_grad(); net.backward()
torch.cuda.synchronize()
o = copy.deepcopy(oracle); o.zero_grad()
rec, mism = {}, {}
l = emulated_forward(o, x, True, taps=plan_taps(net), record=rec, mismatch=mism)
weighted_ce(l, yl, w).backward()
lines = []
for name in reversed(list(rec)):
    t = rec[name]
    act = net.named_acts.get(name)
    if act is None or act.grad is None or t.grad is None:
        continue
    g_plan = act.grad[..., :act.C].permute(0, 3, 1, 2).float()
    g_emu = t.grad
    if act.pre_relu_grad:
        g_emu = g_emu * (t.detach() > 0)
    if name.endswith(".shuf.0.out") or name == "layers.8.0.out":
        from unet_b200.layout import shuffle_row_of_co
        g_plan = g_plan[:, torch.tensor(shuffle_row_of_co(act.C), device=g_plan.device)]
    lines.append(f"{rel(g_plan, g_emu):.3e} cos {cosine(g_plan, g_emu):.5f} fwd {mism.get(name, -1):.2e} pre_relu {int(act.pre_relu_grad)} {name} {tuple(g_plan.shape)}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/grad_bisect.txt", "w").write("\n".join(lines))
print("\n".join(lines[:40]))
