"""Runs N eager (non-graph) training steps of the bench workload — the target of the ncu launch list.
usage: step_eager.py [steps] [batch];  prints launches per step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200.engine import Trainer
from unet_b200.network import UNetB200
from unet_b200.synth import uniform_tiles
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
net = UNetB200("xresnet34", 4, 2, (256, 256), B, training=True)
net.init_parameters(0)
tr = Trainer(net, "sgd", 1e-3, use_graph=False)
x, y = uniform_tiles(B, 4, 256, 256, 2)
x, y = x.cuda(), y.cuda()
torch.cuda.synchronize()
for _ in range(steps):
    tr.step(x, y)
torch.cuda.synchronize()
print("launches_per_step", net.launches_per_train_step)
