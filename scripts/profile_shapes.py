"""Per-shape device time of one steady-state 720p frame (CUDA events around every C-ABI call)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (720, 1280)
net, opt = build_model(prec, torch.device("cuda"))
x = torch.rand(1, 2, 3, H, W, device="cuda")
k = v = None
with torch.no_grad():
    for _ in range(4):
        _, k, v = net(x, k, v)
    eng = net._engine
    eng.profile_begin(shapes=True)
    n = 2
    for _ in range(n):
        _, k, v = net(x, k, v)
    prof = eng.profile_end()
rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
tot = sum(d["ms"] for _, d in rows) / n
print(f"total {tot:.2f} ms/frame")
for name, d in rows:
    ms, L, by, fl = d["ms"] / n, d["launches"] / n, d["bytes"] / n, d["flops"] / n
    print(f"{name:58s} n={L:5.1f} {ms:8.3f} ms  {by/ms/1e6 if ms else 0:8.0f} GB/s  {fl/ms/1e9 if ms else 0:8.1f} TF/s  {by/1e9:7.2f} GB")
