"""Model-level GPU parity: the libb2u.so launch plan vs the fp32 oracle (same weights, same inputs).

north_star tolerances: logits / gradients max|a-b|/max|b| <= 1e-2 in bf16 mode; argmax masks agree >= 99.9 %.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(arch, n_in, n_out, size, batch, seed=0):
    from oracle.unet_oracle import make_oracle
    from unet_b200.network import UNetB200
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    oracle = make_oracle(arch, n_in, n_out, seed=seed).cuda()
    net = UNetB200(arch, n_in, n_out, (size, size), batch, training=True)
    net.load_state_dict(oracle.state_dict())
    g = torch.Generator().manual_seed(1234)
    x_u8 = torch.randint(0, 256, (batch, n_in, size, size), generator=g, dtype=torch.uint8)
    g2 = torch.Generator().manual_seed(4321)
    y = torch.randint(0, n_out, (batch, size, size), generator=g2, dtype=torch.int64)
    return oracle, net, x_u8.cuda(), y.cuda()


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("arch,n_in,n_out,size,batch", [("xresnet34", 4, 2, 64, 2), ("xresnet34", 4, 2, 256, 2),
                                                         ("xresnet18", 3, 2, 128, 4)])
def test_train_step_parity(arch, n_in, n_out, size, batch):
    from oracle.unet_oracle import weighted_ce
    oracle, net, x_u8, y = _setup(arch, n_in, n_out, size, batch)
    oracle.train()
    x = x_u8.float() / 255.0
    w = torch.full((n_out,), 1.0 / n_out, device="cuda")
    logits_ref = oracle(x)
    loss_ref = weighted_ce(logits_ref, y, w)
    loss_ref.backward()

    net.set_input(x_u8)
    net.set_labels(y)
    net.forward()
    loss = net.loss_and_grad()
    net.backward()
    torch.cuda.synchronize()
    logits = net.logits_nchw()
    e_logits = rel(logits, logits_ref)
    agree = (logits.argmax(1) == logits_ref.argmax(1)).float().mean().item()
    e_loss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    report = [f"logits rel {e_logits:.3e} argmax agree {agree:.5f} loss {loss.item():.6f} vs {loss_ref.item():.6f}"]
    worst = 0.0
    grads = net.named_grads()
    for name, p in oracle.named_parameters():
        e = rel(grads[name], p.grad)
        worst = max(worst, e)
        report.append(f"{e:.3e} {name} |ref|max {p.grad.abs().max().item():.3e}")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/parity_{arch}_{size}_{batch}.txt", "w") as f:
        f.write("\n".join(report))
    # running statistics follow torch's update rule
    sd = oracle.state_dict()
    for k, b in net.buffers.items():
        assert rel(b, sd[k]) <= 1e-2, k
    assert e_logits <= 1e-2, report[0]
    assert e_loss <= 1e-2, report[0]
    assert agree >= 0.999, report[0]
    assert worst <= 3e-2, "\n".join(sorted(report[1:], reverse=True)[:12])


def test_eval_forward_parity():
    from oracle.unet_oracle import make_oracle
    from unet_b200.network import UNetB200
    torch.backends.cudnn.allow_tf32 = False
    oracle = make_oracle("xresnet34", 4, 2).cuda().eval()
    net = UNetB200("xresnet34", 4, 2, (256, 256), 2, training=False)
    net.load_state_dict(oracle.state_dict())
    g = torch.Generator().manual_seed(1234)
    x_u8 = torch.randint(0, 256, (2, 4, 256, 256), generator=g, dtype=torch.uint8).cuda()
    with torch.no_grad():
        ref = oracle(x_u8.float() / 255.0)
    net.set_input(x_u8)
    net.forward()
    got = net.logits_nchw()
    torch.cuda.synchronize()
    e = rel(got, ref)
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    assert e <= 1e-2 and agree >= 0.999, (e, agree)
