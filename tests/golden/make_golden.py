"""Generates the committed golden fixtures from the oracle (run in the build container: python tests/golden/make_golden.py).

The reference has no tests or fixtures of its own and cannot be imported here (fastai / slidingwindow /
GDAL are absent), so these vectors pin the ORACLE restatement against drift and give the GPU tests fixed targets;
they do not upgrade the parity status from "unpinned".
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import stitch, windows  # noqa: E402
from oracle.unet_oracle import make_oracle, weighted_ce  # noqa: E402

WINDOW_CASES = [(20000, 20000, 256, 0.125), (1000, 1300, 400, 0.2), (400, 400, 400, 0.0), (257, 511, 256, 0.5),
                (300, 300, 256, 0.75), (64, 64, 256, 0.2)]


def golden_windows():
    out = []
    for h, w, p, ov in WINDOW_CASES:
        ws = windows.compute_windows(h, w, p, ov)
        xs, _ = windows.axis_offsets(w, p, ov)
        ys, _ = windows.axis_offsets(h, p, ov)
        digest = hashlib.sha256(json.dumps(ws).encode()).hexdigest()
        out.append({"height": h, "width": w, "patch": p, "overlap": ov, "count": len(ws), "xs": xs, "ys": ys,
                    "first": ws[:3], "last": ws[-3:], "sha256": digest})
    json.dump(out, open(os.path.join(HERE, "windows.json"), "w"), indent=1)


def golden_stitch():
    h, w, p, ov, c = 70, 90, 32, 0.25, 3
    ws = windows.compute_windows(h, w, p, ov)
    rng = np.random.default_rng(11)
    logits = rng.normal(size=(len(ws), c, p, p)).astype(np.float32) * 2
    probs = [stitch.softmax_probs(l) for l in logits]
    merged = stitch.merge_pixel_windows(probs, ws, h, w)
    gts = [[float(x), float(ww), 1.0, -float(y), float(hh), -1.0] for (x, y, ww, hh) in ws]
    merged8, _ = stitch.merge_tiles(probs, gts, large_file=True)
    np.savez_compressed(os.path.join(HERE, "stitch.npz"), logits=logits, windows=np.array(ws), merged=merged.astype(np.uint8),
                        merged_large_file=merged8.astype(np.uint8), shape=np.array([h, w, p, c]), overlap=np.array([ov]))


def golden_model():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    arch, n_in, n_out, size, batch = "xresnet18", 3, 2, 32, 2
    m = make_oracle(arch, n_in, n_out, seed=0).train()
    g = torch.Generator().manual_seed(1234)
    x_u8 = torch.randint(0, 256, (batch, n_in, size, size), generator=g, dtype=torch.uint8)
    y = torch.randint(0, n_out, (batch, size, size), generator=torch.Generator().manual_seed(4321), dtype=torch.uint8)
    m.eval()                      # eval logits first: the train-mode pass below updates the BN running statistics
    with torch.no_grad():
        logits_eval = m(x_u8.float() / 255.0)
    m.train()
    logits = m(x_u8.float() / 255.0)
    loss = weighted_ce(logits, y.long(), torch.full((n_out,), 1.0 / n_out))
    loss.backward()
    names = ["layers.12.0.weight", "layers.12.0.bias", "layers.11.convpath.1.0.bias", "layers.8.0.0.bias",
             "layers.7.bn.weight", "layers.0.0.0.weight"]
    p = dict(m.named_parameters())
    np.savez_compressed(
        os.path.join(HERE, "model_xresnet18_32.npz"), x_u8=x_u8.numpy(), y=y.numpy(),
        logits_train=logits.detach().numpy(), loss=np.array([loss.item()]), logits_eval=logits_eval.numpy(),
        weight_probe=np.array([p["layers.0.0.0.weight"].detach().flatten()[:8].numpy()]),
        **{"grad::" + n: p[n].grad.numpy() for n in names})


if __name__ == "__main__":
    golden_windows()
    golden_stitch()
    golden_model()
    print(sorted(os.listdir(HERE)))
