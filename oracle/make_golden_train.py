"""Generate tests/golden/train_*.npz from the REFERENCE implementation (TEST INFRASTRUCTURE).

The reference's training step is ``VideoRestorationModel.optimize_parameters`` (VRM:78-108): the cached frame loop
with BPTT through the caches, mean-over-frames L1 loss (Turtle_Derain.yml:105-108), ``torch.optim.AdamW`` built from
the yml's ``optim_g`` (VRM:67-69; lr 4e-4, betas (0.9, 0.99), weight_decay 0).  ``basicsr.models`` cannot be imported
here (matplotlib / lmdb are absent), so this script drives the reference ARCH module (imported by path)
through exactly that loop in fp32 on the CPU (the fp16 autocast + GradScaler of VRM:80,100-105 is a CUDA-only
numerics choice, not part of the algorithm) and records, for two consecutive steps on one batch:

  loss per step, every parameter's gradient at step 1 (as sum / abs-sum digests and verbatim for a few tensors),
  and all parameters after step 2 verbatim.

Usage:  python oracle/make_golden_train.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.make_golden import OUT, load_opt, load_ref_module, tiny_opt  # noqa: E402
from oracle.turtle_oracle import randomize_gates  # noqa: E402

OPTIM = dict(lr=4e-4, weight_decay=0, betas=(0.9, 0.99))             # Turtle_Derain.yml:90-94


def reference_step(net, opt, lq, gt):
    """VRM:78-108 without autocast/GradScaler."""
    opt.zero_grad()
    n = lq.shape[1]
    k = v = None
    l_pix = 0
    for j in range(n):
        pre = lq[:, j if j == 0 else j - 1].unsqueeze(1)
        x = torch.concat([pre, lq[:, j].unsqueeze(1)], dim=1)
        out, k, v = net(x, k, v)
        l_pix = l_pix + torch.nn.functional.l1_loss(out, gt[:, j])
    l_pix = l_pix / n
    total = l_pix + 0 * sum(p.sum() for p in net.parameters())
    total.backward()
    grads = {name: p.grad.detach().clone() for name, p in net.named_parameters()}
    opt.step()
    return float(l_pix), grads


def make_case(variant, frames, H, W, seed):
    opt_yml = tiny_opt(load_opt(variant))
    mod = load_ref_module(variant)
    torch.manual_seed(10)
    net = mod.make_model(opt_yml).train()
    sd = randomize_gates({k: v.detach().clone() for k, v in net.state_dict().items()}, seed=1234)
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    lq = torch.rand(2, frames, 3, H, W, generator=g)                 # batch 2 per GPU, Turtle_Derain.yml:76
    gt = torch.rand(2, frames, 3, H, W, generator=g)
    optim = torch.optim.AdamW([{"params": [p for p in net.parameters() if p.requires_grad]}], **OPTIM)
    loss1, grads = reference_step(net, optim, lq, gt)
    loss2, _ = reference_step(net, optim, lq, gt)
    data = dict(lq=lq.numpy(), gt=gt.numpy(), losses=np.array([loss1, loss2]), variant=np.array(variant),
                opt_yaml=np.array(yaml.safe_dump({k: v for k, v in opt_yml.items() if not isinstance(v, dict)})))
    names = list(grads)
    data["grad_digest"] = np.array([[float(grads[n].double().sum()), float(grads[n].double().abs().sum())]
                                    for n in names])
    data["grad_names"] = np.array(names)
    for n in names:
        data["w::" + n] = sd[n].numpy()
        data["after2::" + n] = net.state_dict()[n].detach().numpy()
    big = sorted(names, key=lambda n: -float(grads[n].abs().sum()))[:6]
    for n in big:
        data["g::" + n] = grads[n].numpy()
    name = f"train_tiny_{variant}.npz"
    np.savez_compressed(os.path.join(OUT, name), **data)
    nz = sum(1 for n in names if float(grads[n].abs().sum()) > 0)
    print(f"[{variant}] losses {loss1:.6f} {loss2:.6f}; {nz}/{len(names)} parameters with a non-zero gradient; wrote "
          f"{name} {os.path.getsize(os.path.join(OUT, name)) // 1024} KiB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    make_case("t0", frames=3, H=64, W=64, seed=21)                   # cfg 5's arch (Turtle_Derain.yml:14)
    make_case("t1", frames=4, H=64, W=64, seed=22)                   # live SAB selection in the graph
