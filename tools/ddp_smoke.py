"""Multi-GPU smoke (run under torchrun): eager DP steps, then CUDA-graph DP steps; prints progress per stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from unet_b200.engine import Trainer, init_distributed
from unet_b200.network import UNetB200
from unet_b200.synth import uniform_tiles

def log(*a):
    print(f"[rank {os.environ.get('RANK')}] ", *a, flush=True)

rank, local, world = init_distributed()
log("dist ok", world)
net = UNetB200("xresnet18", 4, 2, (64, 64), 4, training=True)
net.init_parameters(0)
x, y = uniform_tiles(4, 4, 64, 64, 2, seed=rank)
x, y = x.cuda(), y.cuda()
tr = Trainer(net, "sgd", 1e-3, use_graph=False)
for i in range(3):
    l = tr.step(x, y)
torch.cuda.synchronize()
log("eager steps ok, loss", float(l))
p0 = net.params.clone()
dist.all_reduce(p0, op=dist.ReduceOp.MAX)
log("params identical across ranks:", bool(torch.equal(p0, net.params)))
tr.use_graph = True
t0 = time.time()
for i in range(3):
    l = tr.step(x, y)
torch.cuda.synchronize()
log("graph steps ok, loss", float(l), f"{time.time() - t0:.1f}s")
dist.barrier()
dist.destroy_process_group()
log("done")
