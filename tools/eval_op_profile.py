"""GPU diagnostic: time every op of the EVAL forward plan (prediction path, BN folded) with CUDA events."""
import sys, os, collections, linecache
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops
from unet_b200.network import UNetB200
import unet_b200.network as nw
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = UNetB200("xresnet34", 4, 2, (256, 256), B, training=False)
net.init_parameters(0)
x = torch.randint(0, 256, (B, 4, 256, 256), dtype=torch.uint8, device="cuda")
net.set_input(x)
for _ in range(3):
    net.forward()
torch.cuda.synchronize()
st = torch.cuda.current_stream(); s = ops.stream_ptr()
evs = []
tags = [t for t in net.op_tags if t[0] == "fwd"]
a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record(st)
for op in net.fwd_ops: op(s)
b0.record(st)
for op, tag in zip(net.fwd_ops, tags):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); op(s); b.record(st); evs.append((tag, a, b))
torch.cuda.synchronize()
agg = collections.OrderedDict(); cnt = collections.Counter()
for tag, a, b in evs:
    agg[tag] = agg.get(tag, 0.0) + a.elapsed_time(b); cnt[tag] += 1
tot = sum(agg.values())
out = [f"eval forward, batch {B}: back-to-back {a0.elapsed_time(b0):.2f} ms ({B / a0.elapsed_time(b0) * 1e3:.0f} tiles/s); sum of per-op event times {tot:.2f} ms over {len(evs)} ops"]
for tag, ms in sorted(agg.items(), key=lambda t: -t[1]):
    src = linecache.getline(nw.__file__, tag[2]).strip()[:90]
    out.append(f"{ms:8.3f} ms {100*ms/tot:5.1f}% x{cnt[tag]:3d} {tag[1]}:{tag[2]}  {src}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/eval_op_profile.txt", "w").write("\n".join(out))
print("\n".join(out[:25]))
# per-launch conv table
ops.PROFILE = []
net.forward(); torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
rows = []
for kind, p, a, b in prof:
    d = p.desc; ms = a.elapsed_time(b)
    rows.append((ms, f"{ms:7.3f} ms {p.flops / ms / 1e9:7.1f} TF/s N{d.out.N} {d.out.H}x{d.out.W} Cin{d.w_cin} Cout{d.out.C} taps{d.num_taps} fl{d.flags} res{int(bool(d.res.ptr))}"))
print("conv total", sum(r[0] for r in rows))
for r in sorted(rows, reverse=True)[:14]: print(r[1])
