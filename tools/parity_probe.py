"""Diagnostic (GPU): per-parameter gradient error of the b2u plan vs the fp32 oracle, next to the error of stock
torch bf16 autocast vs the same fp32 oracle (calibrates what bf16 itself costs). Writes gpurun_out/probe_*.txt."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.unet_oracle import make_oracle, weighted_ce
from unet_b200.network import UNetB200
from unet_b200.synth import aerial_like_tiles, uniform_tiles
from oracle.bf16_emulation import emulated_forward

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def run(arch, n_in, n_out, size, batch, data='uniform'):
    oracle = make_oracle(arch, n_in, n_out).cuda().train()
    net = UNetB200(arch, n_in, n_out, (size, size), batch, training=True)
    net.load_state_dict(oracle.state_dict())
    mk = uniform_tiles if data == 'uniform' else aerial_like_tiles
    x_u8, y = mk(batch, n_in, size, size, n_out)
    x_u8, y = x_u8.cuda(), y.cuda().long()
    x = x_u8.float() / 255
    w = torch.full((n_out,), 1.0 / n_out, device="cuda")
    o2 = copy.deepcopy(oracle)
    lr = oracle(x); loss_ref = weighted_ce(lr, y, w); loss_ref.backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l2 = o2(x)
    loss2 = weighted_ce(l2.float(), y, w); loss2.backward()
    o3 = copy.deepcopy(oracle); o3.zero_grad()
    l3 = emulated_forward(o3, x, True); loss3 = weighted_ce(l3, y, w); loss3.backward()
    p3 = dict(o3.named_parameters())
    net.set_input(x_u8); net.set_labels(y); net.forward(); loss = net.loss_and_grad(); net.backward()
    torch.cuda.synchronize()
    # teacher-forced emulation: forward values pinned to the plan's stored activations
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from parity_util import plan_taps, cosine
    o4 = copy.deepcopy(oracle); o4.zero_grad()
    free = {}
    l4 = emulated_forward(o4, x, True, taps=plan_taps(net), record=free); weighted_ce(l4, y, w).backward()
    p4 = dict(o4.named_parameters())
    lg = net.logits_nchw()
    lines = [f"{arch} size {size} batch {batch} data {data}",
             f"vs EMULATION: logits {rel(lg, l3):.3e} loss ours {loss.item():.6f} emu {loss3.item():.6f} argmax agree {(lg.argmax(1) == l3.argmax(1)).float().mean().item():.5f}",
             f"logits: ours {rel(lg, lr):.3e}  autocast {rel(l2, lr):.3e}   loss ours {loss.item():.6f} autocast {loss2.item():.6f} ref {loss_ref.item():.6f}",
             f"argmax agree ours {(lg.argmax(1) == lr.argmax(1)).float().mean().item():.5f} autocast {(l2.argmax(1) == lr.argmax(1)).float().mean().item():.5f}"]
    grads = net.named_grads()
    p2 = dict(o2.named_parameters())
    wo = wa = we = 0
    for name, p in oracle.named_parameters():
        eo, ea, ee = rel(grads[name], p.grad), rel(p2[name].grad, p.grad), rel(grads[name], p3[name].grad)
        wo, wa, we = max(wo, eo), max(wa, ea), max(we, ee)
        et = rel(grads[name], p4[name].grad)
        lines.append(f"{eo:.3e} {ea:.3e} {ee:.3e} TF {et:.3e} cos {cosine(grads[name], p4[name].grad):.5f} {name}")
    lines.insert(4, f"worst grad rel: ours-vs-fp32 {wo:.3e} autocast-vs-fp32 {wa:.3e} ours-vs-emulation {we:.3e}")
    os.makedirs("gpurun_out", exist_ok=True)
    open(f"gpurun_out/probe_{arch}_{size}_{batch}_{data}.txt", "w").write("\n".join(lines))
    print("\n".join(lines[:5]))


if __name__ == "__main__":
    if len(sys.argv) > 1:   # parity_probe.py arch n_in n_out size batch [data]
        a = sys.argv
        run(a[1], int(a[2]), int(a[3]), int(a[4]), int(a[5]), a[6] if len(a) > 6 else 'aerial')
        sys.exit(0)
    for cfg in [("xresnet34", 4, 2, 64, 2, 'uniform'), ("xresnet34", 4, 2, 256, 2, 'uniform'), ("xresnet34", 4, 2, 256, 4, 'aerial'),
                ("xresnet18", 3, 2, 128, 8, 'aerial')]:
        run(*cfg)
