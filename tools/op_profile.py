"""GPU diagnostic: time EVERY op of one eager training step with CUDA events (warm caches, real order) and aggregate by
the network.py source line that registered it."""
import sys, os, collections, linecache
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops
from unet_b200.engine import Trainer
from unet_b200.network import UNetB200
from unet_b200.synth import uniform_tiles
import unet_b200.network as nw

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = UNetB200("xresnet34", 4, 2, (256, 256), B, training=True)
net.init_parameters(0)
tr = Trainer(net, "sgd", 1e-3, use_graph=False)
x, y = uniform_tiles(B, 4, 256, 256, 2)
x, y = x.cuda(), y.cuda()
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
st = torch.cuda.current_stream()
s = ops.stream_ptr()
evs = []
fwd_tags = [t for t in net.op_tags if t[0] == "fwd"]
bwd_tags = [t for t in net.op_tags if t[0] == "bwd"]
def run(ops_list, tags):
    for op, tag in zip(ops_list, tags):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); op(s); b.record(st)
        evs.append((tag, a, b))
net.set_input(tr.x_static, s)
run(net.fwd_ops, fwd_tags)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st); net.loss_and_grad(s); b.record(st); evs.append((("loss", "ce", 0), a, b))
run(net.bwd_ops, bwd_tags)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st); net.sgd_step(1e-3, s); b.record(st); evs.append((("opt", "sgd+stage", 0), a, b))
torch.cuda.synchronize()
agg = collections.OrderedDict(); cnt = collections.Counter()
for tag, a, b in evs:
    agg[tag] = agg.get(tag, 0.0) + a.elapsed_time(b); cnt[tag] += 1
tot = sum(agg.values())
out = [f"batch {B}: sum of per-op event times {tot:.2f} ms over {len(evs)} ops"]
for tag, ms in sorted(agg.items(), key=lambda t: -t[1]):
    src = linecache.getline(nw.__file__, tag[2]).strip()[:90] if tag[2] else ""
    out.append(f"{ms:8.3f} ms {100*ms/tot:5.1f}% x{cnt[tag]:3d} {tag[0]:4s} {tag[1]}:{tag[2]}  {src}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/op_profile.txt", "w").write("\n".join(out))
print("\n".join(out[:40]))
