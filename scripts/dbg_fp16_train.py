import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_training_host import load_train_case
import turtlevsr_b200.training as T
net, lq, gt, z = load_train_case("train_tiny_t1.npz")
net = net.cuda(); lq = lq.cuda(); gt = gt.cuda()
def chk(name, o):
    for i, t in enumerate(o if isinstance(o, (tuple, list)) else [o]):
        if isinstance(t, torch.Tensor) and not torch.isfinite(t.float()).all():
            print("NON-FINITE after", name, i, t.dtype, tuple(t.shape), "nan:", torch.isnan(t).sum().item(), "inf:", torch.isinf(t).sum().item())
            raise SystemExit
for n, m in net.named_modules():
    m.register_forward_hook(lambda m, i, o, n=n: chk(n, o))
for fn in ["_layernorm", "_unit_rows", "_clipped_softmax", "_gated_ffw", "_plain_ffw", "_reduced_attn", "_channel_attn", "_state_align", "_causal_history", "_block"]:
    orig = getattr(T, fn)
    setattr(T, fn, (lambda orig, fn: lambda *a, **k: (lambda r: (chk(fn, r), r)[1])(orig(*a, **k)))(orig, fn))
with torch.autocast("cuda", dtype=torch.float16):
    k = v = None
    for j in range(3):
        pre = lq[:, j if j == 0 else j - 1]
        out, k, v = T.autograd_forward(net, torch.stack([pre, lq[:, j]], 1), k, v)
        print(j, out.dtype, torch.isfinite(out).all().item(), out.abs().max().item())
