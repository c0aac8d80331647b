"""Multi-GPU parity check (run under torch.distributed.run, >= 2 GPUs; tests/test_multigpu.py launches it).

1. Training: every rank holds the same weights and its own batch.  The data-parallel step (CUDA graphs, segment-wise
   all-reduce behind the backward pass, fp32 wire format) must leave on every rank exactly the parameters that rank 0
   computes ALONE by running each rank's batch through a world-size-1 trainer and averaging the gradients itself -
   bit-exact for 2 ranks (a + b == b + a), <= 1e-6 beyond; the bf16 wire format must agree within bf16 rounding.
   (BatchNorm uses per-rank batch statistics, as DDP without SyncBN - SURVEY 8(e).)
2. Prediction: the uint8 mask strips of the ranks, gathered with NCCL, equal the one-GPU mask bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from unet_b200.engine import Trainer, init_distributed
from unet_b200.network import UNetB200
from unet_b200.predict_engine import TiledPredictor, gather_mask_strips
from unet_b200.synth import aerial_like_tiles

rank, local, world = init_distributed()
dev = torch.device("cuda", local)
arch, n_in, n_out, size, B, lr = "xresnet18", 4, 2, 64, 4, 0.05
xs, ys = [], []
for r in range(world):
    x, y = aerial_like_tiles(B, n_in, size, size, n_out, seed=100 + r)
    xs.append(x.to(dev)); ys.append(y.to(dev))


def fresh():
    net = UNetB200(arch, n_in, n_out, (size, size), B, training=True, device=dev)
    net.init_parameters(seed=0)
    return net


ok = True
for wire in ("fp32", "bf16"):
    net = fresh()
    tr = Trainer(net, "sgd", lr, use_graph=True, grad_bf16=(wire == "bf16"))
    tr.step(xs[rank], ys[rank])
    torch.cuda.synchronize()
    got = net.params.clone()
    # reference on this rank alone: every rank's gradient from a world-size-1 plan, averaged here
    ref_net = fresh()
    gsum = torch.zeros_like(ref_net.grads)
    for r in range(world):
        ref_net.set_input(xs[r]); ref_net.set_labels(ys[r]); ref_net.forward(); ref_net.loss_and_grad(); ref_net.backward()
        torch.cuda.synchronize()
        g = ref_net.grads.clone()
        gsum += g.to(torch.bfloat16).float() if wire == "bf16" else g
    # the same optimizer kernel on the summed gradients (identical arithmetic: lr * (1/world) * g, one FMA per element)
    ref_net.grads.copy_(gsum)
    from unet_b200 import _lib, ops
    _lib.check(ref_net.lib.b2u_sgd_step(ref_net.params.data_ptr(), ref_net.grads.data_ptr(), ref_net.layout.total, lr,
                                        1.0 / world, ops.stream_ptr()), "b2u_sgd_step")
    torch.cuda.synchronize()
    want = ref_net.params
    err = ((got - want).abs().max() / want.abs().max()).item()
    same = torch.equal(got, want)
    tol = 0.0 if (wire == "fp32" and world == 2) else (1e-6 if wire == "fp32" else 2e-4)
    good = same if tol == 0.0 else err <= tol
    # every rank ends with the same parameters
    pm = got.clone(); dist.all_reduce(pm, op=dist.ReduceOp.MAX)
    good = good and torch.equal(pm, got)
    ok = ok and good
    if rank == 0:
        print(f"ddp_check train wire={wire} world={world}: max rel err {err:.3e} bit_exact={same} identical_on_all_ranks={torch.equal(pm, got)} -> {'OK' if good else 'FAIL'}", flush=True)
    del tr, net, ref_net

# ---- prediction: gathered strips == one-GPU mask
ev = UNetB200(arch, n_in, n_out, (size, size), 8, training=False, device=dev)
ev.init_parameters(seed=3)
pred = TiledPredictor(ev)
raster, _ = aerial_like_tiles(1, n_in, 300, 520, n_out, seed=5)
raster = raster[0].to(dev).contiguous()
strip, xb, xe = pred.predict_raster(raster, 0.125, rank, world)
full = gather_mask_strips(strip, raster.shape[2], rank, world)
one, _, _ = pred.predict_raster(raster, 0.125, 0, 1)
from unet_b200.predict_engine import gather_mask_cells
grid = (1, world)          # row cells: the other axis of the 2-D ownership grid
cell, rect = pred.predict_raster(raster, 0.125, rank, world, grid=grid)
full2 = gather_mask_cells(cell, raster.shape[1], raster.shape[2], grid, rank, world)
if rank == 0:
    same = torch.equal(full, one) and torch.equal(full2, one)
    ok = ok and same
    print(f"ddp_check predict world={world}: gathered mask == one-GPU mask: {same} -> {'OK' if same else 'FAIL'}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
