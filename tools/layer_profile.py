"""GPU diagnostic: per-launch table of the GEMM kernels in one eager training step (CUDA events on the stream)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops
from unet_b200.engine import Trainer
from unet_b200.network import UNetB200
from unet_b200.synth import uniform_tiles

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = UNetB200("xresnet34", 4, 2, (256, 256), B, training=True)
net.init_parameters(0)
tr = Trainer(net, "sgd", 1e-3, use_graph=False)
x, y = uniform_tiles(B, 4, 256, 256, 2)
x, y = x.cuda(), y.cuda()
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
# whole-step eager timing per phase
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
st = torch.cuda.current_stream()
s = ops.stream_ptr()
tr.x_static.copy_(x); net.labels.copy_(y)
ev[0].record(st); net.set_input(tr.x_static, s); net.forward(s); ev[1].record(st)
net.loss_and_grad(s); ev[2].record(st); net.backward(s); ev[3].record(st)
net.sgd_step(1e-3, s); ev[4].record(st)
torch.cuda.synchronize()
print(f"eager phases ms: fwd {ev[0].elapsed_time(ev[1]):.2f} loss {ev[1].elapsed_time(ev[2]):.2f} bwd {ev[2].elapsed_time(ev[3]):.2f} opt+stage {ev[3].elapsed_time(ev[4]):.2f}")
ops.PROFILE = []
tr.step(x, y)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
rows = []
for kind, p, a, b in prof:
    ms = a.elapsed_time(b)
    d = p.desc
    if kind == "conv":
        shape = f"N{d.out.N} {d.out.H}x{d.out.W} Cin{d.w_cin} Cout{d.out.C} taps{d.num_taps} BN{p.info.block_n} mt{p.info.m_tiles} nt{p.info.n_tiles} st{p.info.stages} grid{p.info.grid} fl{d.flags} res{int(bool(d.res.ptr))} zm{int(bool(d.zmask.ptr))}"
    else:
        shape = f"N{d.dy.N} {d.dy.H}x{d.dy.W} Cin{d.Cin} Cout{d.Cout} taps{d.num_taps} BN{p.info.block_n} T{p.info.taps_per_unit} splits{p.info.splits} units{p.info.units} ks{p.info.k_steps} st{p.info.stages} bias{d.want_bias}"
    rows.append((kind, ms, p.flops / (ms * 1e-3) / 1e12, shape))
tot = {}
for k, ms, tf, sh in rows:
    tot[k] = tot.get(k, 0) + ms
out = [f"batch {B}; totals ms: {tot}"]
for k, ms, tf, sh in rows:
    out.append(f"{k:5s} {ms:8.3f} ms {tf:7.1f} TF/s  {sh}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/layer_profile.txt", "w").write("\n".join(out))
print(out[0])
