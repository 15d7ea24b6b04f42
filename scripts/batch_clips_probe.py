"""Probe: frames/s when B independent 720p clips share one batched forward per frame step (CUDA graphs on)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
dev = torch.device("cuda")
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    net, _ = build_model("tf32", dev)
    net.enable_cuda_graphs()
    g = torch.Generator().manual_seed(0)
    clip = torch.rand(4, B, 3, 720, 1280, generator=g).to(dev)
    k = v = None
    n_warm, n = 20, 12
    with torch.no_grad():
        for j in range(n_warm + n):
            if j == n_warm:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            x = torch.stack([clip[(j - 1) % 4 if j else 0], clip[j % 4]], 1)
            _, k, v = net(x, k, v)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"B={B}: {dt * 1e3 / n:.2f} ms/step, {B * n / dt:.2f} frames/s, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del net, k, v, clip
    torch.cuda.empty_cache()
