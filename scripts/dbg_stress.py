"""Localise a/b divergence of the ring stress test: compare workspace buffers after every frame."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from helpers import load_case
from turtlevsr_b200.archs import create_video_model

def build(opt, sd, precision):
    net = create_video_model(opt); net.load_state_dict(sd, strict=True)
    return net.cuda().eval().set_precision(precision)

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 40
opt, sd, _, _, _ = load_case("tiny_super_live.npz")
for rep in range(3):
    g = torch.Generator().manual_seed(44)
    clip = torch.rand(1, nfr, 3, 24, 32, generator=g).cuda()
    a = build(opt, sd, prec); b = build(opt, sd, prec)
    ka = va = kb = vb = None
    ndiff = 0
    for j in range(nfr):
        x = torch.stack([clip[:, max(j - 1, 0)], clip[:, j]], 1)
        oa, ka, va = a(x, ka, va)
        ob, kb, vb = b(x, kb, vb)
        kb = [None if t is None else t.clone() for t in kb]
        vb = [None if t is None else t.clone() for t in vb]
        if not torch.equal(oa, ob):
            ndiff += 1
            if ndiff <= 2:
                print(f"rep {rep} frame {j}: out max|d| = {(oa-ob).abs().max().item():.3e}")
                for name in a._engine.ws.bufs:
                    ta, tb = a._engine.ws.bufs[name], b._engine.ws.bufs.get(name)
                    if tb is None or ta.shape != tb.shape: print("   ", name, "shape differs"); continue
                    if not torch.equal(ta, tb):
                        d = (ta.float() - tb.float()).abs()
                        print(f"    {name:12s} differs: n={int((d>0).sum())}/{d.numel()} max={d.max().item():.3e}")
                for i in range(8):
                    if ka[i] is not None:
                        print("    cache", i, "k equal", torch.equal(ka[i], kb[i]), "v equal", torch.equal(va[i], vb[i]))
    print(f"rep {rep}: {ndiff} differing frames of {nfr}")
