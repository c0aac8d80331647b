"""BASELINE configs[2]: tiled prediction of a synthetic 4-band raster (256x256 tiles, 32-px overlap), stitched argmax
mask.  usage: predict_bench.py [side=20000] [batch=64]   (torchrun for N > 1: tiles sharded by output column strips)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from unet_b200.engine import init_distributed
from unet_b200.network import UNetB200
from unet_b200.predict_engine import TiledPredictor, gather_mask_strips
from unet_b200.tiling import compute_windows

side = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rank, local, world = init_distributed()
dev = torch.device("cuda", local)
net = UNetB200("xresnet34", 4, 2, (256, 256), B, training=False, device=dev)
net.init_parameters(0)
# counter-hash raster generated on the device block by block (no 1.6 GB host file)
g = torch.Generator(device=dev).manual_seed(1234)
raster = torch.randint(0, 256, (4, side, side), dtype=torch.uint8, device=dev, generator=g)
pred = TiledPredictor(net)
small = raster[:, :1024, :1024].contiguous()
pred.predict_raster(small, 0.125)          # warm-up
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
mask, xb, xe = pred.predict_raster(raster, 0.125, rank, world)
torch.cuda.synchronize()
t_pred = time.perf_counter() - t0
full = gather_mask_strips(mask, side, rank, world)      # the only collective of the path: uint8 strips -> rank 0
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t)
n_tiles = len(compute_windows(side, side, 256, 0.125))
if rank == 0:
    print(json.dumps({"workload": f"predict {side}x{side} 4-band raster, 256 tiles / 32 px overlap", "n_gpus": world, "tiles": n_tiles,
                      "tiles_run_rank0": pred.tiles_run, "seconds": dt, "seconds_before_gather": t_pred,
                      "tiles_per_s": n_tiles / dt, "mask_shape_rank0": list(full.shape),
                      "mask_class1_fraction": float((full == 1).float().mean())}))
if world > 1: dist.destroy_process_group()
