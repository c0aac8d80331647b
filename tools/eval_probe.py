import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.unet_oracle import make_oracle
from unet_b200.network import UNetB200
from unet_b200.synth import aerial_like_tiles
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()
ARCH, NIN, SZ = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("xresnet34", 4, 256)
oracle = make_oracle(ARCH, NIN, 2).cuda().eval()
net = UNetB200(ARCH, NIN, 2, (SZ, SZ), 2, training=False)
net.load_state_dict(oracle.state_dict())
x_u8, _ = aerial_like_tiles(2, NIN, SZ, SZ, 2); x_u8 = x_u8.cuda()
acts = {}
enc = oracle.layers[0]
for i, child in enumerate(enc):
    child.register_forward_hook(lambda m, inp, out, i=i: acts.__setitem__(i, out))
with torch.no_grad(): ref = oracle(x_u8.float() / 255)
net.set_input(x_u8); net.forward(); torch.cuda.synchronize()
names = {}
for n_, m in oracle.named_modules():
    if n_.startswith("layers.0.") and n_.count(".") == 3 and n_.split(".")[2] in "4567":
        m.register_forward_hook(lambda mod, inp, out, n_=n_: names.__setitem__(n_ + ".out", out))
    if n_ in ("layers.0.0", "layers.0.1", "layers.0.2"):
        m.register_forward_hook(lambda mod, inp, out, n_=n_: names.__setitem__(n_ + ".out", out))
with torch.no_grad(): ref = oracle(x_u8.float() / 255)
for k, v in names.items():
    if k in net.named_acts:
        a = net.named_acts[k]
        print(k, rel(a.t[..., :a.C].permute(0, 3, 1, 2), v))
for k, a in net.feats.items():
    print("feat", k, rel(a.t[..., :a.C].permute(0, 3, 1, 2), acts[k]))
dec = {}
for i in range(4, 8):
    oracle.layers[i].register_forward_hook(lambda m, inp, out, i=i: dec.__setitem__(f"layers.{i}.conv2.out", out))
oracle.layers[3].register_forward_hook(lambda m, inp, out: dec.__setitem__("layers.3.1.out", out))
oracle.layers[2].register_forward_hook(lambda m, inp, out: dec.__setitem__("enc.bnrelu", out))
with torch.no_grad(): ref = oracle(x_u8.float() / 255)
for k, v in dec.items():
    a = net.named_acts[k]
    print(k, rel(a.t[..., :a.C].permute(0, 3, 1, 2), v))
print("logits", rel(net.logits_nchw(), ref))
